"""Diagnostic (GPU box): is the reference's never-initialised d_laplace (cudaMalloc, fd-code.cu:176;
quirk Q2) really zero?  Probe it in a fresh process, before and after another process used the GPU."""
import ctypes as C, os, sys, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref as R

def probe(tag):
    g = R.GpuFam()
    nxe, nze = 149, 131
    g.fd_init(8, nxe, nze, 24, 24, 10, 1, 0.75, 10.0, 10.0, 0.001)
    ptr = C.c_void_p.in_dll(g.L, "d_laplace").value
    rt = C.CDLL("libcudart.so.12")
    host = np.empty(nxe * nze, np.float32)
    rc = rt.cudaMemcpy(host.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), C.c_size_t(host.nbytes), 2)
    nz_ = np.count_nonzero(host.view(np.uint32))
    print("%s: cudaMemcpy rc=%d, d_laplace non-zero words: %d of %d, max|finite| %g" % (
        tag, rc, nz_, host.size, np.nanmax(np.abs(np.where(np.isfinite(host), host, 0)))), flush=True)

if len(sys.argv) > 1:
    probe(sys.argv[1])
else:
    subprocess.call([sys.executable, __file__, "fresh process #1"])
    # dirty the GPU memory from another process, then free it
    subprocess.call([sys.executable, "-c", "import torch; x=torch.randn(1<<28, device='cuda'); y=torch.randn(1<<28, device='cuda'); torch.cuda.synchronize(); print('dirtied 2 GiB')"])
    subprocess.call([sys.executable, __file__, "fresh process #2 (after another process wrote 2 GiB)"])
    subprocess.call([sys.executable, __file__, "fresh process #3"])
