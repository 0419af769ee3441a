import ctypes as C, os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
from diffdart_b200 import _abi, skeletons
lib=_abi.load_library(); N,n=65536,5
topo,params,keep=skeletons.create_multiarm_robot(n).to_ctypes()
h=_abi._vp(); assert lib.ddb_create(C.byref(topo),C.byref(params),N,0,C.byref(h))==0
lib.ddb_set_lanes(h, int(os.environ.get("LANES","2")))
rng=np.random.default_rng(0)
hin=[torch.tensor(rng.uniform(-1,1,(N,n))).pin_memory() for _ in range(5)]
hout=[torch.empty(N,n,dtype=torch.float64).pin_memory() for _ in range(5)]
def step():
    assert lib.ddb_forward_host(h,*[t.data_ptr() for t in hin[:3]],hout[0].data_ptr(),hout[1].data_ptr())==0
    assert lib.ddb_backward_host(h,hin[3].data_ptr(),hin[4].data_ptr(),*[t.data_ptr() for t in hout[2:]])==0
for _ in range(5): step()
t0=time.perf_counter()
for _ in range(50): step()
dt=(time.perf_counter()-t0)/50
print("chunks",os.environ.get("DDB_HOST_CHUNKS"),"lanes",os.environ.get("LANES"),"ms/step %.3f"%(dt*1e3),"ws/s %.3e"%(N/dt))
# zero-copy: the kernels read the pinned host buffers and write the pinned host results themselves (UVA), no staging copies
st = torch.cuda.current_stream(); sp = C.c_void_p(st.cuda_stream)
def zstep():
    assert lib.ddb_forward(h, sp, *[t.data_ptr() for t in hin[:3]], hout[0].data_ptr(), hout[1].data_ptr(), None, None, None, None) == 0
    st.synchronize()
    assert lib.ddb_backward(h, sp, hin[3].data_ptr(), hin[4].data_ptr(), *[t.data_ptr() for t in hout[2:]]) == 0
    st.synchronize()
ref = [t.clone() for t in hout]
for t in hout: t.zero_()
for _ in range(5): zstep()
t0=time.perf_counter()
for _ in range(50): zstep()
dt=(time.perf_counter()-t0)/50
print("zero-copy ms/step %.3f"%(dt*1e3),"ws/s %.3e"%(N/dt), "equal", all(torch.equal(a,b) for a,b in zip(ref,hout)))
# hybrid: inputs by bulk H2D copy, results written to host memory by the kernel
dbuf=[torch.empty(N,n,dtype=torch.float64,device="cuda") for _ in range(5)]
def hstep():
    for a,b in zip(dbuf[:3],hin[:3]): a.copy_(b,non_blocking=True)
    assert lib.ddb_forward(h, sp, *[t.data_ptr() for t in dbuf[:3]], hout[0].data_ptr(), hout[1].data_ptr(), None, None, None, None) == 0
    st.synchronize()
    for a,b in zip(dbuf[3:],hin[3:]): a.copy_(b,non_blocking=True)
    assert lib.ddb_backward(h, sp, dbuf[3].data_ptr(), dbuf[4].data_ptr(), *[t.data_ptr() for t in hout[2:]]) == 0
    st.synchronize()
for _ in range(5): hstep()
t0=time.perf_counter()
for _ in range(50): hstep()
dt=(time.perf_counter()-t0)/50
print("hybrid (H2D copy in, kernel writes host) ms/step %.3f"%(dt*1e3),"ws/s %.3e"%(N/dt), "equal", all(torch.equal(a,b) for a,b in zip(ref,hout)))
