import os, sys, time, ctypes as C
import numpy as np, torch
sys.path[:0]=['.','drl-tetris_b200','tests','oracle']
import bench
from b200tetris import Engine
from b200tetris.engine import _hp
n=1<<20
dev=torch.device('cuda:0')
# raw copy bandwidth
h=torch.empty(23*1024*1024,dtype=torch.uint8).pin_memory(); d=torch.empty_like(h,device=dev)
for _ in range(3): d.copy_(h,non_blocking=True)
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(20): d.copy_(h,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/20
print("H2D 23MiB pinned: %.3f ms  %.1f GB/s"%(dt*1e3, h.numel()/dt/1e9))
h4=torch.empty(4*1024*1024,dtype=torch.uint8).pin_memory(); d4=torch.empty_like(h4,device=dev)
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(20): h4.copy_(d4,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/20
print("D2H 4MiB pinned: %.3f ms  %.1f GB/s"%(dt*1e3, h4.numel()/dt/1e9))
from b200tetris import host_array
for wc in (False, True):
    hb=host_array(23*1024*1024,np.uint8,write_combined=wc); hb[:]=1
    hbt=torch.from_numpy(hb)
    for _ in range(3): d.copy_(hbt,non_blocking=True)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(20): d.copy_(hbt,non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/20
    print("H2D 23MiB b2t_host_alloc wc=%s: %.3f ms  %.1f GB/s"%(wc,dt*1e3,hb.size/dt/1e9))
WC=os.environ.get("E2E_WC","0")=="1"
eng=Engine(n)
gids=np.arange(n)
eng.reset((bench.splitmix64(0xB200,gids,0)&np.uint64(0xFFFF)).astype(np.uint32))
def pinned(a):
    if WC:
        h=host_array(a.shape,a.dtype,write_combined=True); h[...]=a; return h
    return torch.from_numpy(a).pin_memory().numpy()
hk=[]
for t in range(4):
    k,ln=bench.sventon_keys(gids,t)
    hk.append((pinned(k),pinned(ln),pinned((bench.splitmix64(0xB200,gids,t+1)&np.uint64(0xFFFF)).astype(np.uint32)),pinned(((gids+t)&1).astype(np.uint8))))
_wc=WC; WC=False
h_rew=pinned(np.empty(n,np.int8)); h_done=pinned(np.empty(n,np.uint8)); h_lines=pinned(np.empty((n,2),np.uint8))
WC=_wc
MODE=os.environ.get("E2E_MODE","lean")
hp4=[pinned(np.ascontiguousarray(k[:,0::2]|(k[:,1::2]<<4))) for k,_,_,_ in hk]
def host_step(t):
    k,ln,sd,pl=hk[t%4]
    keys,lens,seeds,fmt,auto=k,None,None,0,1
    if MODE=="explicit": lens,seeds,auto=ln,sd,0
    if MODE=="packed": keys,fmt=hp4[t%4],1
    rc=eng.lib.b2t_step_host(eng._h,None,n,_hp(pl,C.c_uint8),_hp(keys,C.c_uint8),_hp(lens,C.c_uint8),k.shape[1],None,1,0,_hp(seeds,C.c_uint32),_hp(h_rew,C.c_int8),_hp(h_done,C.c_uint8),_hp(h_lines,C.c_uint8),None,None,None,None,fmt,auto)
    assert rc==0
for t in range(3): host_step(t)
t0=time.perf_counter()
for t in range(10): host_step(t)
dt=(time.perf_counter()-t0)/10
print("wc=%s mode=%s slices=%s  step_host %.3f ms  %.3e env-steps/s"%(WC,MODE,os.environ.get("B2T_PIPE_SLICES","default"),dt*1e3,n/dt))
