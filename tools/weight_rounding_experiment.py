#!/usr/bin/env python3
"""Which rounding sets the Gram error of the fp16 forward path, and what does error-feedback rounding of the weights buy?
CPU emulation against an fp64 run of the same network (random-init VGG-19, seed 1234, 1/f^2 images): per style layer the
relative Frobenius error of the Gram matrix with (a) weights and activations rounded to fp16 (round to nearest), (b) weights
only, (c) activations only, (d) weights rounded with the error of each weight carried into the next weight of the same filter
(torch order: input channel major, then taps) + activations, (e) the same carried across input channels within a tap only,
(f) (d) with exact activations.  Output: profiles/r02_gram_error_by_rounding_source.log.  CPU only; uses the oracle."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nst_oracle as O
import torch.nn.functional as F
torch.set_num_threads(4)
ws,bs=O.vgg19_random_weights(1234,13)
def h16(t): return t.half().float()
def diffuse(w):
    # error-diffusion rounding to fp16 along the K dimension (cin*9) per output channel: keeps sum of rounding errors ~0
    co=w.shape[0]; flat=w.reshape(co,-1).double()
    out=torch.empty_like(flat)
    err=torch.zeros(co,dtype=torch.float64)
    for k in range(flat.shape[1]):
        v=flat[:,k]+err
        r=v.float().half().double()
        # candidate neighbours
        out[:,k]=r
        err=v-r
    return out.float().reshape(w.shape)
def diffuse_tap(w):
    # diffusion over cin within each (cout, tap)
    co,ci,kh,kw=w.shape
    x=w.permute(0,2,3,1).reshape(co*9,ci).double()
    out=torch.empty_like(x); err=torch.zeros(co*9,dtype=torch.float64)
    for k in range(ci):
        v=x[:,k]+err; r=v.float().half().double(); out[:,k]=r; err=v-r
    return out.float().reshape(co,kh,kw,ci).permute(0,3,1,2).contiguous()
def fwd(x, W, act_round, dtype=torch.float32):
    taps={}
    conv=0
    x=x.to(dtype)
    for item in O.VGG_CFG:
        if item=="M": x=F.max_pool2d(x,2,2); continue
        x=F.conv2d(x,W[conv].to(dtype),bs[conv].to(dtype),padding=1)
        name=O.CONV_NAMES[conv]
        if name in O.STYLE_LAYERS: taps[name]=act_round(x) if conv>0 else x
        if conv==12: break
        x=F.relu(x)
        if conv>0 or True: x=act_round(x) if conv>0 else act_round(x)
        conv+=1
    return taps
def gram(x):
    b,c,h,w=x.shape; f=x.reshape(c,h*w).double(); return f@f.t()/(c*h*w)
for (H,Wd,seed) in [(64,64,0),(50,38,3),(128,128,0)]:
    img=O.to_tensor_u8(O.synth_image(H,Wd,seed)); x=O.normalize(img,O.VGG_MEAN,O.VGG_STD)
    ref=fwd(x,[w.double() for w in ws],lambda t:t,torch.float64)
    W_rn=[ws[0]]+[h16(w) for w in ws[1:]]
    W_df=[ws[0]]+[diffuse(w) for w in ws[1:]]
    W_dt=[ws[0]]+[diffuse_tap(w) for w in ws[1:]]
    res={}
    for tag,W,ar in (("rn w+act",W_rn,h16),("rn w only",W_rn,lambda t:t),("act only",ws,h16),("diffuse w+act",W_df,h16),("diffuse-tap w+act",W_dt,h16),("diffuse w only",W_df,lambda t:t)):
        t=fwd(x,W,ar)
        res[tag]=[float((gram(t[n])-gram(ref[n])).norm()/gram(ref[n]).norm()) for n in O.STYLE_LAYERS]
    print(H,Wd)
    for k,v in res.items(): print("  %-18s"%k," ".join("%.2e"%e for e in v))
