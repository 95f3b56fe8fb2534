"""utils/global_step_functions.py against the known answers of utils/global_step_functions_test.py:25-42."""
import pytest

from tensor2robot_b200.utils import global_step_functions as gsf


@pytest.mark.parametrize('boundaries,values,test_inputs,expected', [
    ([1], [5.0], [0, 1, 10], [5.0, 5.0, 5.0]),                                              # constant
    ([10, 20], [1.0, 11.0], [0, 10, 13, 15, 18, 20, 25], [1.0, 1.0, 4.0, 6.0, 9.0, 11.0, 11.0]),   # ramp_up
])
def test_piecewise_linear_known_answers(boundaries, values, test_inputs, expected):
  fn = gsf.piecewise_linear(boundaries, values)
  assert [fn(x) for x in test_inputs] == expected


def test_piecewise_linear_asserts_and_exponential_decay():
  with pytest.raises(AssertionError):
    gsf.piecewise_linear([], [])
  with pytest.raises(AssertionError):
    gsf.piecewise_linear([0, 1], [1.0])
  stair = gsf.exponential_decay(0.1, decay_steps=10, decay_rate=0.5, staircase=True)
  assert [stair(s) for s in (0, 9, 10, 25)] == [0.1, 0.1, 0.05, 0.025]
  smooth = gsf.exponential_decay(0.1, decay_steps=10, decay_rate=0.5, staircase=False)
  assert smooth(5) == pytest.approx(0.1 * 0.5 ** 0.5)
  # usable as an optimizer learning rate
  from tensor2robot_b200.models import optimizers
  assert optimizers.MomentumOptimizer(stair, 0.9).learning_rate(10) == 0.05
