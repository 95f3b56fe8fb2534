"""TEST INFRASTRUCTURE ONLY - evaluates a TensorFlow-written GraphDef with numpy, op by op, without TensorFlow.

Used to pin the CPU restatements in oracle/ against a graph that TensorFlow itself serialised: the reference's fixture
`test_data/mock_exported_savedmodel/saved_model.pb` (copied to tests/golden/mock_saved_model.pb) is the exported
inference graph of utils/mocks.py MockT2RModel.  Only the protobuf wire format (SavedModel -> MetaGraphDef -> GraphDef ->
NodeDef, AttrValue, TensorProto) and the handful of ops that graph uses are implemented; variables are read from the
TF-written checkpoint beside it."""
import struct

import numpy as np

_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_}


def _varint(b, i):
  r = s = 0
  while True:
    c = b[i]
    i += 1
    r |= (c & 0x7F) << s
    s += 7
    if c < 0x80:
      return r, i


def _fields(b):
  i, out = 0, []
  while i < len(b):
    key, i = _varint(b, i)
    f, w = key >> 3, key & 7
    if w == 0:
      v, i = _varint(b, i)
    elif w == 2:
      n, i = _varint(b, i)
      v = bytes(b[i:i + n])
      i += n
    elif w == 5:
      v = bytes(b[i:i + 4])
      i += 4
    elif w == 1:
      v = bytes(b[i:i + 8])
      i += 8
    else:
      raise ValueError('wire type %d' % w)
    out.append((f, w, v))
  return out


def _tensor(proto):
  """TensorProto -> numpy (dtype = 1, tensor_shape = 2, tensor_content = 4, float_val = 5, int_val = 7, int64_val = 10)."""
  dtype, shape, content, vals = np.float32, [], None, []
  for f, w, v in _fields(proto):
    if f == 1:
      dtype = _DTYPES[v]
    elif f == 2:
      for f2, _, dim in _fields(v):
        if f2 == 2:
          size = [x for g, _, x in _fields(dim) if g == 1]
          shape.append(size[0] if size else 0)
    elif f == 4:
      content = v
    elif f == 5:
      vals += list(struct.unpack('<%df' % (len(v) // 4), v)) if w == 2 else [struct.unpack('<f', v)[0]]
    elif f in (7, 10):
      if w == 2:
        i = 0
        while i < len(v):
          x, i = _varint(v, i)
          vals.append(x - (1 << 64) if x >> 63 else x)
      else:
        vals.append(v - (1 << 64) if v >> 63 else v)
  if content is not None:
    return np.frombuffer(content, dtype=dtype).reshape(shape).copy()
  n = int(np.prod(shape)) if shape else 1
  if len(vals) == 1 and n > 1:
    vals = vals * n
  return np.asarray(vals, dtype=dtype).reshape(shape)


class Graph(object):
  """nodes: {name: (op, inputs, attrs)}; attrs hold raw AttrValue bytes."""

  def __init__(self, saved_model_bytes):
    meta = [v for f, _, v in _fields(saved_model_bytes) if f == 2][0]
    graph_def = [v for f, _, v in _fields(meta) if f == 2][0]
    self.nodes = {}
    for f, _, node in _fields(graph_def):
      if f != 1:
        continue
      name = op = None
      inputs, attrs = [], {}
      for g, _, v in _fields(node):
        if g == 1:
          name = v.decode()
        elif g == 2:
          op = v.decode()
        elif g == 3:
          inputs.append(v.decode())
        elif g == 5:
          kv = dict((k, x) for k, _, x in _fields(v))
          attrs[kv[1].decode()] = kv.get(2, b'')
      self.nodes[name] = (op, inputs, attrs)

  def ops_on_path(self, output):
    """Op types of the compute nodes `output` depends on, in evaluation order (variables / constants left out)."""
    seen, order = set(), []

    def visit(name):
      name = name.split(':')[0].lstrip('^')
      if name in seen:
        return
      seen.add(name)
      op, inputs, _ = self.nodes[name]
      if op in ('VariableV2', 'VarHandleOp', 'Const', 'Placeholder'):
        return
      for i in inputs:
        visit(i)
      if op not in ('Identity', 'ReadVariableOp'):
        order.append(op)
    visit(output)
    return order

  def run(self, output, feeds, variables):
    cache = {}

    def attr_bool(attrs, key):
      return any(f == 5 and v for f, _, v in _fields(attrs.get(key, b'')))

    def ev(name):
      name = name.split(':')[0]
      if name in cache:
        return cache[name]
      op, inputs, attrs = self.nodes[name]
      if name in feeds:
        r = np.asarray(feeds[name])
      elif op in ('VariableV2', 'VarHandleOp'):
        r = np.asarray(variables[name])
      elif op in ('Identity', 'ReadVariableOp'):
        r = ev(inputs[0])
      elif op == 'Const':
        r = _tensor([v for f, _, v in _fields(attrs['value']) if f == 8][0])
      elif op == 'MatMul':
        a, b = ev(inputs[0]), ev(inputs[1])
        r = (a.T if attr_bool(attrs, 'transpose_a') else a) @ (b.T if attr_bool(attrs, 'transpose_b') else b)
      elif op in ('BiasAdd', 'AddV2', 'Add'):
        r = ev(inputs[0]) + ev(inputs[1])
      elif op == 'Sub':
        r = ev(inputs[0]) - ev(inputs[1])
      elif op == 'Mul':
        r = ev(inputs[0]) * ev(inputs[1])
      elif op == 'Rsqrt':
        r = 1.0 / np.sqrt(ev(inputs[0]))
      elif op == 'Elu':
        x = ev(inputs[0])
        r = np.where(x > 0, x, np.expm1(np.minimum(x, 0)))
      else:
        raise NotImplementedError('op %s (%s)' % (op, name))
      cache[name] = r
      return r

    return ev(output)
