import ctypes as C, os, sys, time
import numpy as np
ROOT='/root/repo'
sys.path.insert(0, ROOT)
import _d1pkg
pkg=_d1pkg.load_pkg()
from dav1d_mirror_b200 import frame as F
L=pkg.lib()
n=64
hfs=[F.HostFrame(3840,2160,0x3ff,1000+i) for i in range(4)]
for hf in hfs:
    d=hf.intra.reshape(-1,40)
    m=d[:,15]
    sel=(m<=14)
    d[sel,15]=0; d[sel,16]=0
    hf.record_levels()
ctx=F.open_context(0)
dfs=[]
for s in range(n):
    hf=hfs[s%4]; df=F.DeviceFrame(ctx,hf); df.upload_descriptors()
    for r in range(2): df.upload_picture(df.refs[r], F.random_planes(hf,7+r))
    df.upload_picture(df.dst, F.random_planes(hf,99)); dfs.append(df)
mf=F.MultiFrame(ctx,dfs,phase_mask=16)
e0,e1=L.dav1d_cuda_event_create(),L.dav1d_cuda_event_create()
for _ in range(2): mf.launch()
L.dav1d_cuda_synchronize(ctx)
L.dav1d_cuda_event_record(ctx,e0)
for _ in range(3): mf.launch()
L.dav1d_cuda_event_record(ctx,e1)
print("all-DC intra class: %.1f us/frame" % (L.dav1d_cuda_event_elapsed_ms(e0,e1)/(3*n)*1e3))
pkg.check_error()
