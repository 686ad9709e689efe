"""Constants of the reference that the hot path reads (reference: src/config.py:4-34).

Only values are mirrored; the reference module's side effects (creating ./logs and ./dataset_cache at
import time, src/config.py:41-45) are deliberately not reproduced.
"""
RANDOM_SEED = 42
NUM_JOINTS = 17
BATCH_SIZE = 10
GRADIENT_ACCUMULATION_STEPS = 10

# loss weights (src/config.py:15-18)
INTER_JOINT_LOSS_WEIGHT = 100
ABS_ROOT_LOSS_WEIGHT = 1
L1_LOSS_WEIGHT = 1
MSE_LOSS_WEIGHT = 1

# optimiser (src/config.py:21-22)
LEARNING_RATE = 0.001
WEIGHT_DECAY = 0.01

# augmentation ranges (src/config.py:28-34)
USE_AUGMENTATION = False
ROTATION_RANGE = (-30, 30)
FLIP_PROB = 0.5
SCALE_RANGE = (0.8, 1.2)
TRANSLATE_RANGE = (-0.1, 0.1)
BRIGHTNESS_RANGE = (0.8, 1.2)
CONTRAST_RANGE = (0.8, 1.2)

# H36M left/right joint pairs swapped by a horizontal flip (src/dataset/augmentation.py:226-234)
SYMMETRIC_JOINTS = [(1, 4), (2, 5), (3, 6), (11, 14), (12, 15), (13, 16)]
