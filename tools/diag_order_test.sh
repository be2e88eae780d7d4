set -e
cd /root/repo
python - <<'PY'
import os, sys, subprocess, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import test_gpu_drivers as T, parity_cases as PC
from oracle import ref as R
import tempfile
nx,nz,nb,nt,ns=101,83,24,400,2
base=tempfile.mkdtemp()
imgs={}
for name,prog in [("ref_a",R.path("rtm_code_ref")),("ref_b",R.path("rtm_code_ref")),("ours",os.path.join(T.BIN,"rtm_code")),("ref_c",R.path("rtm_code_ref")),("ref_d",R.path("rtm_code_ref"))]:
    d=os.path.join(base,name); os.makedirs(d); T._write_rtm_case(d,nx,nz,nb,nt,ns,seed=3)
    subprocess.run([prog,"./input.dat"],cwd=d,capture_output=True,check=True)
    imgs[name]=np.fromfile(os.path.join(d,"out","dir.image"),np.float32)
for a in imgs:
    print(a, " ".join("%s:%.2e"%(b,PC.rel_l2(imgs[a],imgs[b])) for b in imgs))
PY
