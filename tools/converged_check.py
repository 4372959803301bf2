import os, sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle import orc
from ptsharp_b200 import bindings
import importlib.util
spec=importlib.util.spec_from_file_location('tgp','/root/repo/tests/test_gpu_parity.py'); m=importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
dev=bindings.Device(0)
for name,(W,H) in (("c1",(48,36)),("c2",(40,40)),("c3",(48,27))):
    for seed in (123, 7):
        hw,ow,_=m._worlds(orc,bindings,name); dev.upload(hw)
        spp,passes=8,24
        ref,var,_=ow.render(W,H,spp,passes=passes,threads=os.cpu_count(),rng_mode=orc.RNG_SEQUENTIAL,seed=seed)
        dev.reset_buffer()
        for i in range(passes): dev.render_pass(hw.make_pass(W,H,spp,pass_index=100+i+seed),want_mean=False)
        img=dev.read_buffer(W,H,0).astype(np.float64); gvar=dev.read_buffer(W,H,1).astype(np.float64); dev.reset_buffer()
        lr,l=ref.mean(axis=2),img.mean(axis=2)
        sigma=np.sqrt((var.mean(axis=2)+gvar.mean(axis=2))/passes)
        z=np.abs(l-lr)/(sigma+0.01*lr+1e-3)
        print(name,seed,'mean_rel',abs(l.mean()-lr.mean())/lr.mean(),'z>4',(z>4).sum(),'of',z.size,'zmax',z.max(),'ppr',np.abs(l-lr).mean()/lr.mean())
