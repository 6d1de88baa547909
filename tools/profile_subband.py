"""One forward + inverse sub-band transform of a 16384 x 16384 x 3 image at the depth given on the command line
(for `ncu --set full -k regex:forward|inverse`); prints nothing but the two launch times."""
import ctypes as C
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from wicca_b200 import _capi
from wicca_b200.plan import pitch_bytes

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 3
lib = _capi.load()
dev = torch.device("cuda:0")
S = 16384
pitch = pitch_bytes(S, 3)
img = torch.randint(0, 256, (S, pitch), dtype=torch.uint8, device=dev)
coeffs = torch.empty((S, S, 3), dtype=torch.float32, device=dev)
work = torch.empty((S * S * 3 * 5 // 16 + 64,), dtype=torch.float32, device=dev)
rec = torch.empty((S, S, 3), dtype=torch.float32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    _capi.check(lib.wicca_haar_forward_dev(img.data_ptr(), S, S, 3, pitch, depth, 1, 0.0, coeffs.data_ptr(), work.data_ptr(), 0,
                                           C.c_void_p(stream)), "forward_dev")
    _capi.check(lib.wicca_haar_inverse_dev(coeffs.data_ptr(), S, S, 3, depth, rec.data_ptr(), work.data_ptr(), 0,
                                           C.c_void_p(stream)), "inverse_dev")
torch.cuda.synchronize()
print("ok depth", depth)
