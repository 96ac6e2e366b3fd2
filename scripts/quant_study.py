import sys; sys.path.insert(0,'.')
import numpy as np
from oracle import embedding as E, mel, heads as H
from openwakeword_b200 import weights as W
emb = W.synthetic_embedding(0)
rng = np.random.default_rng(5)
wins=[]
for i in range(48):
    amp=[300,3000,12000][i%3]
    x=np.clip(rng.normal(0,amp,12400+512),-32768,32767).astype(np.int16)
    wins.append(mel.melspectrogram(x)[:76])
wins=np.stack(wins).astype(np.float32)

def fwd(qa_from, qw_from, qa_to=20, qw_to=20):
    """quantise activations INTO layers li in [qa_from, qa_to) and weights of layers [qw_from,qw_to) to fp16"""
    x = wins[...,None].astype(np.float64)
    for li,(kh,kw,cin,cout,pool) in enumerate(E.LAYERS):
        w = emb["conv"][li].astype(np.float64); a = x
        if li>0 and qa_from<=li<qa_to: a = a.astype(np.float16).astype(np.float64)
        if li>0 and qw_from<=li<qw_to: w = w.astype(np.float16).astype(np.float64)
        x = E._conv(a,w,np.float64)
        if li==0: x=np.maximum(x,0)
        if li<19:
            s,b=E.fold_bn(*[np.asarray(p,dtype=np.float64) for p in emb["bn"][li]])
            x=x*s.astype(np.float64)+b.astype(np.float64)
            x=np.maximum(float(E.LEAK)*x,x); x=np.maximum(x,float(E.FLOOR))
        if pool is not None: x=E._pool(x,*pool)
    return x[:,0,0,:]
ref = fwd(99,99)
for name,args in [("all fp16",(1,1)),("act only",(1,99)),("w only",(99,1)),
                  ("fp16 layers <19",(1,1,19,19)),("fp16 layers <18",(1,1,18,18)),("fp16 layers <17",(1,1,17,17)),
                  ("fp16 layers <15",(1,1,15,15)),("fp16 layers <11",(1,1,11,11)),("fp16 layers <7",(1,1,7,7))]:
    y = fwd(*args); e=np.abs(y-ref)
    print(f"{name:20s} emb max err {e.max():.3e} mean {e.mean():.3e}")
