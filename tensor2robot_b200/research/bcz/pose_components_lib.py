"""Action / state component tuples of BC-Z (research/bcz/pose_components_lib.py)."""
# Name, size, whether it is residual or not, and loss weight: parameterises action labels.
# Name, size, whether residual or not: parameterises proprioceptive state inputs.
DEFAULT_STATE_COMPONENTS = []
DEFAULT_ACTION_COMPONENTS = [
    ('xyz', 3, True, 100.),
    ('quaternion', 4, False, 10.),
    ('target_close', 1, False, 1.),
]
