"""Image encoding helpers of the replay-writing side (utils/image.py:25-60)."""
import io

import numpy as np
from PIL import Image
from PIL import ImageFile


def jpeg_string(image, jpeg_quality=90):
  """A PIL image as JPEG bytes (quality 1..95, optimised Huffman tables)."""
  ImageFile.MAXBLOCK = 640 * 512 * 64     # large frames need a larger encoder buffer
  out = io.BytesIO()
  image.save(out, 'jpeg', quality=jpeg_quality, optimize=True)
  return out.getvalue()


def numpy_to_image_string(image_array, image_format='jpeg', data_type=np.uint8):
  """A numpy image as an encoded image string."""
  out = io.BytesIO()
  Image.fromarray(image_array.astype(data_type)).save(out, image_format)
  return out.getvalue()
