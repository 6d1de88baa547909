"""wicca_b200 - the WICCA HaarCoder hot path on NVIDIA B200 (sm_100a).

Public surface (mirrors ``wicca/wavelet_coder.py`` of the reference):

    from wicca_b200 import HaarCoder, WaveletCoder
    icon = HaarCoder().get_small_copy(image, transform_depth)

The arithmetic lives in ``libwicca_b200.so`` (hand-written CUDA behind the C ABI of
``include/wicca_b200.h``); there is no CPU fallback.
"""
from .data_loader import UnsupportedImageError, decode_jpeg, icons_from_jpeg, icons_from_jpeg_files, jpeg_info, load_image
from .wavelets import CoifletCoder, DaubechiesCoder, OrthogonalWaveletCoder
from .wavelet_coder import (BORDER_CONSTANT, BORDER_REFLECT, BORDER_REFLECT_101, BORDER_REPLICATE, BORDER_WRAP,
                            HaarCoder, WaveletCoder, list_to_mallat, mallat_to_list, validate_image)

__all__ = ["HaarCoder", "WaveletCoder", "validate_image", "mallat_to_list", "list_to_mallat",
           "OrthogonalWaveletCoder", "DaubechiesCoder", "CoifletCoder",
           "load_image", "decode_jpeg", "jpeg_info", "icons_from_jpeg", "icons_from_jpeg_files", "UnsupportedImageError",
           "BORDER_CONSTANT", "BORDER_REPLICATE", "BORDER_REFLECT", "BORDER_WRAP", "BORDER_REFLECT_101"]
__version__ = "0.1.0"
