#!/bin/bash
# AddressSanitizer fuzzing of the host-side parsers (JPEG headers + Huffman, TFRecord index, tf.Example /
# SequenceExample): writes seed inputs, builds tests/native/fuzz/host_fuzz.cc with -fsanitize=address,undefined and runs it.
#   scripts/host_fuzz.sh [iterations]        (default 200000; no GPU needed)
set -euo pipefail
cd "$(dirname "$0")/.."
ITER=${1:-200000}
SEEDS=$(mktemp -d)
python - "$SEEDS" <<'PY'
import io, os, sys
import numpy as np
from PIL import Image
sys.path.insert(0, '.')
from oracle import tfrecord
out = sys.argv[1]
rng = np.random.RandomState(0)
a = rng.randint(0, 256, (48, 64, 3)).astype(np.uint8)
for name, kw in (('444', dict(subsampling=0)), ('422', dict(subsampling=1)), ('420', dict(subsampling=2)),
                 ('rst', dict(subsampling=2, restart_marker_blocks=2))):
  buf = io.BytesIO(); Image.fromarray(a).save(buf, format='JPEG', quality=80, **kw)
  open(os.path.join(out, name + '.jpg'), 'wb').write(buf.getvalue())
buf = io.BytesIO(); Image.fromarray(a[..., 0]).save(buf, format='JPEG')
open(os.path.join(out, 'grey.jpg'), 'wb').write(buf.getvalue())
records = tfrecord.read_tfrecords('tests/golden/pose_env_test_data.tfrecord')[:4]
tfrecord.write_tfrecords(os.path.join(out, 'examples.tfrecord'), records)
extra = [tfrecord.make_example({'pose': [0.5, -1.0], 'reward': [1.0], 'ids': [1, -2, 3], 'state/image': b'xyz'}) for _ in range(3)]
tfrecord.write_tfrecords(os.path.join(out, 'examples2.tfrecord'), extra)
seqs = [tfrecord.make_sequence_example({'reward': [1.0]}, {'pose': [[0.1 * t, 0.2 * t] for t in range(n)]}) for n in (1, 3, 5)]
tfrecord.write_tfrecords(os.path.join(out, 'seq.tfrecord'), seqs)
PY
g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fno-sanitize-recover=all -fwrapv -std=c++17 -pthread \
    tests/native/fuzz/host_fuzz.cc tensor2robot_b200/csrc/jpeg_host.cc tensor2robot_b200/csrc/host_io.cc -o "$SEEDS/host_fuzz"
"$SEEDS/host_fuzz" "$SEEDS" "$ITER"
rm -rf "$SEEDS"
