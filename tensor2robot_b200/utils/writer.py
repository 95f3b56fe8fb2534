"""TFRecordReplayWriter: saves transitions to a TFRecord-backed replay buffer (utils/writer.py:27-61).  The framing is
TFRecord's: uint64 length, masked CRC-32C of the length, payload, masked CRC-32C of the payload; the CRC is the
library's host routine (`t2r_masked_crc32c`, the one the reader verifies with)."""
import ctypes as C
import os
import struct

from tensor2robot_b200 import _lib


def _masked_crc(data):
  buf = (C.c_uint8 * len(data)).from_buffer_copy(data) if data else None
  return int(_lib.lib().t2r_masked_crc32c(buf, len(data))) & 0xFFFFFFFF


def frame_record(payload):
  header = struct.pack('<Q', len(payload))
  return header + struct.pack('<I', _masked_crc(header)) + payload + struct.pack('<I', _masked_crc(payload))


class TFRecordReplayWriter(object):
  """open(path) -> write(list of Examples)* -> close(); the file is `path + '.tfrecord'`."""

  def __init__(self):
    self.writer = None

  def open(self, path):
    if self.writer is not None:
      raise ValueError('Writer is already open!')
    dirname = os.path.dirname(path)
    if dirname and not os.path.isdir(dirname):
      os.makedirs(dirname)
    self.writer = open(path + '.tfrecord', 'wb')

  def close(self):
    if self.writer is None:
      raise ValueError('Writer is not open!')
    self.writer.close()
    self.writer = None

  def write(self, transitions):
    """Writes an entire episode: `transitions` is a list of Example-like objects with SerializeToString()."""
    if self.writer is None:
      raise ValueError('Writer is not open!')
    for transition in transitions:
      self.writer.write(frame_record(transition.SerializeToString()))
