"""Message classes of proto/t2r.proto, built at import time without protoc.

Wire- and text-format compatible with the reference's proto/t2r.proto:18-43 (same package, message
and field names/numbers), so `assets.extra/t2r_assets.pbtxt` files are interchangeable.
"""
from google.protobuf import descriptor_pb2
from google.protobuf import descriptor_pool
from google.protobuf import message_factory

_PACKAGE = 'third_party.py.tensor2robot'
_F = descriptor_pb2.FieldDescriptorProto


def _field(msg, name, number, ftype, label=_F.LABEL_OPTIONAL, type_name=None):
  f = msg.field.add()
  f.name, f.number, f.type, f.label = name, number, ftype, label
  if type_name:
    f.type_name = type_name
  return f


def _build():
  fdp = descriptor_pb2.FileDescriptorProto()
  fdp.name = 'tensor2robot_b200/proto/t2r.proto'
  fdp.package = _PACKAGE
  fdp.syntax = 'proto2'

  spec = fdp.message_type.add()
  spec.name = 'ExtendedTensorSpec'
  _field(spec, 'shape', 1, _F.TYPE_INT32, _F.LABEL_REPEATED)
  _field(spec, 'dtype', 2, _F.TYPE_INT32)
  _field(spec, 'name', 3, _F.TYPE_STRING)
  _field(spec, 'is_optional', 4, _F.TYPE_BOOL)
  _field(spec, 'is_extracted', 5, _F.TYPE_BOOL)
  _field(spec, 'data_format', 6, _F.TYPE_STRING)
  _field(spec, 'dataset_key', 7, _F.TYPE_STRING)
  _field(spec, 'varlen_default_value', 8, _F.TYPE_FLOAT)

  struct = fdp.message_type.add()
  struct.name = 'TensorSpecStruct'
  entry = struct.nested_type.add()
  entry.name = 'KeyValueEntry'
  entry.options.map_entry = True
  _field(entry, 'key', 1, _F.TYPE_STRING)
  _field(entry, 'value', 2, _F.TYPE_MESSAGE, type_name='.%s.ExtendedTensorSpec' % _PACKAGE)
  _field(struct, 'key_value', 1, _F.TYPE_MESSAGE, _F.LABEL_REPEATED,
         '.%s.TensorSpecStruct.KeyValueEntry' % _PACKAGE)

  assets = fdp.message_type.add()
  assets.name = 'T2RAssets'
  _field(assets, 'feature_spec', 1, _F.TYPE_MESSAGE, type_name='.%s.TensorSpecStruct' % _PACKAGE)
  _field(assets, 'label_spec', 2, _F.TYPE_MESSAGE, type_name='.%s.TensorSpecStruct' % _PACKAGE)
  _field(assets, 'global_step', 3, _F.TYPE_INT32)

  pool = descriptor_pool.DescriptorPool()
  pool.Add(fdp)
  get = getattr(message_factory, 'GetMessageClass', None)
  out = {}
  for name in ('ExtendedTensorSpec', 'TensorSpecStruct', 'T2RAssets'):
    desc = pool.FindMessageTypeByName('%s.%s' % (_PACKAGE, name))
    out[name] = get(desc) if get else message_factory.MessageFactory(pool).GetPrototype(desc)
  return out


_classes = _build()
ExtendedTensorSpec = _classes['ExtendedTensorSpec']
TensorSpecStruct = _classes['TensorSpecStruct']
T2RAssets = _classes['T2RAssets']
