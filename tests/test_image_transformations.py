"""Crops and flips of preprocessors/image_transformations.py with the cases of the reference's
image_transformations_test.py:67-300 (ramp images: R = x coordinate, G = B = y coordinate).  These ops are index work on
tensors of any device; the photometric kernels are covered by tests/test_ops_parity_gpu.py."""
import numpy as np
import pytest
import torch

from tensor2robot_b200.preprocessors import image_transformations as it


def _ramp(batch_size, height, width):
  mesh_x, mesh_y = np.meshgrid(np.arange(width, dtype=np.float32), np.arange(height, dtype=np.float32))
  image = np.stack([mesh_x, mesh_y, mesh_y], 2)
  return torch.from_numpy(np.tile(image[None], (batch_size, 1, 1, 1)))


@pytest.mark.parametrize('output_shape', [[20, 20], [32, 32]])
def test_random_crop(output_shape):
  images = _ramp(4, 32, 32)
  cropped = it.RandomCropImages([images], [32, 32, 3], output_shape)[0].numpy()
  assert list(cropped.shape) == [4] + output_shape + [3]
  assert cropped[0, -1, 0, 1] - cropped[0, 0, 0, 1] == output_shape[0] - 1
  assert cropped[0, 0, -1, 0] - cropped[0, 0, 0, 0] == output_shape[1] - 1


def test_random_crop_draws_one_offset_for_the_whole_list():
  it.seed(0)
  a, b = it.RandomCropImages([_ramp(2, 32, 32), _ramp(2, 32, 32) + 100], [32, 32, 3], [8, 8])
  np.testing.assert_array_equal(a.numpy() + 100, b.numpy())
  offsets = set()
  for _ in range(40):
    c = it.RandomCropImages([_ramp(1, 32, 32)], [32, 32, 3], [8, 8])[0]
    offsets.add((int(c[0, 0, 0, 1]), int(c[0, 0, 0, 0])))
  assert len(offsets) > 10 and all(0 <= y <= 24 and 0 <= x <= 24 for y, x in offsets)


@pytest.mark.parametrize('crop_fn', [it.RandomCropImages, it.CenterCropImages])
def test_wrong_crop_arguments(crop_fn):
  images = _ramp(4, 32, 32)
  with pytest.raises(ValueError):
    crop_fn([images], [32, 32, 3], [20, 64])             # larger than the input (testFaultyRandomCrop)
  with pytest.raises(ValueError):
    crop_fn([images], [32, 32], [20, 64])                # input shape must be (height, width, channels)
  with pytest.raises(ValueError):
    crop_fn([images], [32, 32, 3, 4], [20, 64])
  with pytest.raises(ValueError):
    crop_fn([images], [32, 32, 3], [20])                 # target shape must be (height, width)
  with pytest.raises(ValueError):
    crop_fn([images], [32, 32, 3], [20, 32, 64])


@pytest.mark.parametrize('input_shape,output_shape', [([32, 32], [20, 20]), ([512, 640], [472, 472])])
def test_center_crop(input_shape, output_shape):
  images = _ramp(4, input_shape[0], input_shape[1])
  cropped = it.CenterCropImages([images], input_shape + [3], output_shape)[0].numpy()
  assert list(cropped.shape) == [4] + output_shape + [3]
  assert cropped[0, 0, 0, 1] == (input_shape[0] - output_shape[0]) // 2
  assert cropped[0, -1, 0, 1] == (input_shape[0] - output_shape[0]) // 2 + output_shape[0] - 1
  assert cropped[0, 0, 0, 0] == (input_shape[1] - output_shape[1]) // 2
  assert cropped[0, 0, -1, 0] == (input_shape[1] - output_shape[1]) // 2 + output_shape[1] - 1


@pytest.mark.parametrize('target_shape', [[20, 20], [32, 32]])
def test_custom_crop(target_shape):
  images = _ramp(4, 32, 32)
  target_locations = np.tile(np.array([[10, 10]]), [4, 1])
  cropped = it.CustomCropImages([images], [32, 32, 3], target_shape, [target_locations])[0].numpy()
  assert list(cropped.shape) == [4] + target_shape + [3]
  assert cropped[0, -1, 0, 1] - cropped[0, 0, 0, 1] == target_shape[0] - 1
  assert cropped[0, 0, -1, 0] - cropped[0, 0, 0, 0] == target_shape[1] - 1


def test_custom_crop_windows_follow_each_batch_element():
  images = _ramp(3, 32, 40)
  locations = np.array([[16, 20], [0, 0], [31, 39]])              # centre, clamped to the top-left / bottom-right
  cropped = it.CustomCropImages([images], [32, 40, 3], [8, 10], [locations])[0].numpy()
  assert [(int(c[0, 0, 1]), int(c[0, 0, 0])) for c in cropped] == [(12, 15), (0, 0), (24, 30)]
  shared = it.CustomCropImages([images], [32, 40, 3], [8, 10], [np.array([16, 20])])[0]
  np.testing.assert_array_equal(shared[1].numpy(), cropped[0])


def test_faulty_custom_crop():
  images = _ramp(4, 32, 32)
  locations = np.tile(np.array([[10, 10]]), [4, 1])
  with pytest.raises(ValueError):
    it.CustomCropImages([images], [32, 32, 3], [53, 8], [locations])       # testFaultyCustomCrop
  with pytest.raises(ValueError):
    it.CustomCropImages([images, images], [32, 32, 3], [8, 8], [locations])
  with pytest.raises(ValueError):
    it.CustomCropImages([images], [32, 32, 3], [8, 8], [locations[:2]])


def test_random_flips_apply_to_the_whole_batch():
  images = _ramp(2, 4, 6)
  seen = set()
  it.seed(3)
  for _ in range(60):
    flipped = it.ApplyRandomFlips(images).numpy()
    lr = bool(flipped[0, 0, 0, 0] == 5)
    ud = bool(flipped[0, 0, 0, 1] == 3)
    np.testing.assert_array_equal(flipped[0], flipped[1])          # consistent across the batch
    seen.add((lr, ud))
  assert seen == {(False, False), (False, True), (True, False), (True, True)}
