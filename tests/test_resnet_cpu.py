"""The parts of layers/resnet_test.py:26-69 that need no device: a malformed FiLM block mask is refused before any
kernel runs, and the engine refuses CPU tensors loudly (no CPU fallback)."""
import functools

import pytest
import torch

from tensor2robot_b200 import _lib
from tensor2robot_b200 import nn
from tensor2robot_b200.layers import resnet


def test_malformed_film_raises():
  image = torch.zeros((2, 224, 224, 3), dtype=torch.bfloat16)
  embedding = torch.zeros((2, 100), dtype=torch.float32)
  film_generator_fn = functools.partial(resnet.linear_film_generator, enabled_block_layers=[True] * 5)
  with nn.variable_store(nn.VariableStore('cpu', seed=0)):
    with pytest.raises(ValueError):
      resnet.resnet_model(image, is_training=True, num_classes=1001, resnet_size=18, return_intermediate_values=True,
                          film_generator_fn=film_generator_fn, film_generator_input=embedding)


def test_no_cpu_fallback():
  image = torch.zeros((2, 64, 64, 3), dtype=torch.bfloat16)
  with nn.variable_store(nn.VariableStore('cpu', seed=0)):
    with pytest.raises(_lib.T2RError, match='no CPU path'):
      resnet.resnet_model(image, is_training=False, num_classes=10, resnet_size=18)
  with pytest.raises(_lib.T2RError, match='no CPU path'):
    nn.relu(torch.zeros(4))
  with pytest.raises(_lib.T2RError, match='no CPU path'):
    nn.layer_norm(torch.zeros(2, 4))
