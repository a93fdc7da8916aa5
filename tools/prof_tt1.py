import sys,os
sys.path.insert(0,"tt-sketch_b200"); sys.path.insert(0,".")
import numpy as np, torch, time
from tt_sketch.drm import TensorTrainDRM
from tt_sketch.sketch import stream_sketch
from tt_sketch.tensor import TensorTrain
shape=(10000,10000,10000,500); lr,rr=(20,)*3,(40,)*3
L=TensorTrainDRM(lr,shape=shape,transpose=False,seed=1); R=TensorTrainDRM(rr,shape=shape,transpose=True,seed=2)
t=TensorTrain.random(shape,10,seed=5)
for _ in range(2): stream_sketch(t,lr,rr,left_drm=L,right_drm=R)
torch.cuda.synchronize()
