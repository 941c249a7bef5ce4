import ctypes as C, os, sys
import torch
here = os.path.dirname(os.path.abspath(__file__))
lib = C.CDLL(os.path.join(here, "libexp_halo.so"))
lib.exp_halo.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
torch.manual_seed(0)
H = W = 24
x = torch.randn(1, H, W, 64).bfloat16().cuda()
w = (torch.randn(64, 64) * 0.2).bfloat16()
wswz = torch.empty(64, 64, dtype=torch.bfloat16)
for r in range(64):
    for j in range(8):
        wswz[r, (j ^ (r & 7)) * 8:(j ^ (r & 7)) * 8 + 8] = w[r, j * 8:j * 8 + 8]
wswz = wswz.cuda()
xf, wf = x.float().cpu()[0], w.float()
for boxw in (16, 10, 8):
    for kh in (0, 1, 2):
        for kw in (0, 1, 2):
            if boxw == 8 and kw != 0:
                continue
            exp = torch.zeros(128, 64)
            for ty in range(16):
                for tx in range(8):
                    yy, xx = ty + kh - 1, tx + kw - 1
                    if 0 <= yy < H and 0 <= xx < W:
                        exp[ty * 8 + tx] = wf @ xf[yy, xx]
            for bo in sorted({0, kw, (kw + 2 * kh) & 7, 8 - kw if kw else 0}):
                out = torch.full((128, 64), float("nan"), device="cuda")
                rc = lib.exp_halo(x.data_ptr(), H, W, wswz.data_ptr(), out.data_ptr(), boxw, kh, kw, bo & 7)
                err = float((out.cpu() - exp).abs().max()) if rc == 0 else -1
                print("boxw %2d kh %d kw %d base_offset %d -> rc %d max_err %.4f %s" % (boxw, kh, kw, bo & 7, rc, err, "OK" if 0 <= err < 0.05 else ""), flush=True)
