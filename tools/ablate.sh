#!/bin/bash
# timing ablations of the elementwise stages (results are numerically wrong by construction; timing only)
# The ablation libraries are built in the dev container first (nvcc cross-compiles; build/ travels with gpurun):
#   for a in 0 1 2 3; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared \
#       -DFA_ABLATE=$a -o build/libfa_ablate$a.so flash_attention_dlrs_b200/csrc/*.cu; done
mkdir -p gpurun_out build
for a in 0 1 2 3; do
  [ -f build/libfa_ablate$a.so ] || { echo "build/libfa_ablate$a.so missing (see the header of this script)"; continue; }
  FA_B200_LIB=build/libfa_ablate$a.so python - <<PY
import os, torch, sys
sys.path.insert(0, os.getcwd())
from flash_attention_dlrs_b200 import _native
B,H,N,D=2,32,8192,128
dev=torch.device("cuda",0)
g=torch.Generator().manual_seed(42)
Q,K,V,dO=(torch.randn(B,H,N,D,generator=g).to(torch.bfloat16).to(dev) for _ in range(4))
sc=D**-0.5
O,L=_native.forward(Q,K,V,True,sc)
delta=_native.backward_preprocess(O,dO)
def t(fn,reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/reps
print("ablate $a  fwd %.3f  dkdv %.3f  dq %.3f ms" % (
  t(lambda:_native.forward(Q,K,V,True,sc)),
  t(lambda:_native.backward(Q,K,V,O,dO,L,True,sc,1,delta)),
  t(lambda:_native.backward(Q,K,V,O,dO,L,True,sc,2,delta))))
PY
done 2>&1 | tee gpurun_out/ablate.log
