"""Golden values for the reference fixture test_data/pose_env_test_data.tfrecord (copied verbatim
to tests/golden/pose_env_test_data.tfrecord), produced with the pure-Python oracle reader + PIL.

  python tests/golden/make_pose_env_golden.py
"""
import io
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import tfrecord  # noqa: E402

records = tfrecord.read_tfrecords(os.path.join(HERE, 'pose_env_test_data.tfrecord'))
pose, reward, target, img_sum, img_first, lengths = [], [], [], [], [], []
for rec in records:
  ex = tfrecord.parse_example(rec)
  assert sorted(ex) == ['pose', 'reward', 'state/image', 'target_pose'], sorted(ex)
  pose.append(ex['pose'][1])
  reward.append(ex['reward'][1])
  target.append(ex['target_pose'][1])
  img = np.asarray(Image.open(io.BytesIO(ex['state/image'][1][0])).convert('RGB'))
  assert img.shape == (64, 64, 3)
  img_sum.append(int(img.astype(np.int64).sum()))
  img_first.append(img[:2, :2].copy())
  lengths.append(len(rec))
np.savez_compressed(os.path.join(HERE, 'pose_env_golden.npz'), pose=np.array(pose, np.float32),
                    reward=np.array(reward, np.float32), target_pose=np.array(target, np.float32),
                    image_sum=np.array(img_sum), image_corner=np.array(img_first), record_length=np.array(lengths))
print('records', len(records), 'lengths', min(lengths), max(lengths))
