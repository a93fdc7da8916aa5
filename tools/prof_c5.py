import os, sys, time
import numpy as np
ROOT="/root/repo"
for p in (ROOT, os.path.join(ROOT,"tt-sketch_b200")): sys.path.insert(0,p)
import torch
from tt_sketch import _backend as be
from tt_sketch.drm import TensorTrainDRM, SparseGaussianDRM
from tt_sketch.sketch import stream_sketch
from tt_sketch.tensor import SparseTensor, TensorSum, TensorTrain
shape=(10000,10000,10000,500)
lr,rr=(20,)*3,(40,)*3
L=TensorTrainDRM(lr,shape=shape,transpose=False,seed=1); R=TensorTrainDRM(rr,shape=shape,transpose=True,seed=2)
def timed(fn,reps=3):
    fn(); torch.cuda.synchronize(); best=1e9
    for _ in range(reps):
        t0=time.perf_counter(); fn(); torch.cuda.synchronize(); best=min(best,time.perf_counter()-t0)
    return best
for nnz in (2_000_000, 20_000_000):
    idx=np.stack([np.random.default_rng(200+k).integers(0,n,nnz) for k,n in enumerate(shape)]).astype(np.int64)
    sp=SparseTensor(shape,idx,np.random.default_rng(99).standard_normal(nnz))
    l0=be.launch_count()
    t=timed(lambda: stream_sketch(sp,lr,rr,left_drm=L,right_drm=R))
    print(f"sparse TT-DRM nnz={nnz}: {t*1e3:.2f} ms  {nnz/t:.3e} nnz/s launches/call={(be.launch_count()-l0)//4}")
    from ctypes import byref, c_double
    a,b=c_double(),c_double(); be.lib().ttsk_last_kernel_ms(be.ctx(),byref(a),byref(b)); print("   kernels ms total/pass",a.value,b.value)
tts=[TensorTrain.random(shape,10,seed=1000+k) for k in range(20)]
t=timed(lambda: stream_sketch(TensorSum(tts),lr,rr,left_drm=L,right_drm=R))
print(f"20 TT summands: {t*1e3:.2f} ms")
t=timed(lambda: stream_sketch(tts[0],lr,rr,left_drm=L,right_drm=R))
print(f"1 TT summand: {t*1e3:.2f} ms")
