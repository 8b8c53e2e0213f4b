"""AddressSanitizer + UBSan over the product HOST layer (cf_host.c) linked with tests/stub_device/stub_device.c (CPU stand-in for the
device ABI): LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tools/asan_host_layer.py"""
import ctypes as C, os, subprocess, sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import harness as H
from corrfunc_b200 import _capi
out="/tmp/libhoststub_asan.so"
host=os.path.join(H.ROOT,"corrfunc_b200","csrc","host")
subprocess.check_call(["/usr/bin/gcc","-std=c11","-O1","-g","-fsanitize=address,undefined","-fno-omit-frame-pointer","-fPIC","-shared","-ffp-contract=off","-fopenmp",
    "-I",os.path.join(H.ROOT,"include"),"-I",host,os.path.join(host,"cf_host.c"),os.path.join(H.ROOT,"tests","stub_device","stub_device.c"),"-o",out,"-lm"])
lib=C.CDLL(out)
ra,dec,cz,w=H.load_mr19_mock_cz()
o=_capi.default_options(np.float64,bin_refine_factors=(1,1,1))
r=_capi.call_vpf_mocks(lib,10.0,10,10000,6,1,H.VPF_CENTERS,1,ra[:20000],dec[:20000],cz[:20000],options=o); print("vpf_mocks file ok")
for dtype in (np.float64,np.float32):
    g,dg,d,_=H.mock_points(41,5000,dtype); rr,rd,rdd,_=H.mock_points(42,1500,dtype)
    o=_capi.default_options(dtype,bin_refine_factors=(1,1,1),is_comoving_dist=True)
    r=_capi.call_vpf_mocks(lib,12.0,6,60,4,2,"/tmp/asan_centres.txt",1,g,dg,d,RAND_RA=rr,RAND_DEC=rd,RAND_CZ=rdd,options=o); print("vpf_mocks randoms ok",dtype.__name__)
    x,y,z,_=H.box_points(9,5000,300.0,dtype)
    for per in (True,False):
        o=_capi.default_options(dtype,periodic=per,boxsize=300.0 if per else None,bin_refine_factors=(1,1,1))
        r=_capi.call_vpf(lib,12.0,6,200,5,77,x,y,z,options=o)
    print("vpf theory ok",dtype.__name__)
    cz5=(d*dtype(60.0)).astype(dtype)
    o=_capi.default_options(dtype,need_avg_sep=True,is_comoving_dist=False)
    r=_capi.call_DDrppi_mocks(lib,1,1,1,20.0,np.logspace(0,1.3,6),g,dg,cz5,w1=np.ones_like(g),weight_type="pair_product",options=o); print("rppi mocks cz ok",dtype.__name__)
    o=_capi.default_options(dtype,need_avg_sep=True,is_comoving_dist=True)
    r=_capi.call_DDsmu_mocks(lib,0,1,1,0.9,5,np.logspace(0,1.3,6),g,dg,d,RA2=rr,DEC2=rd,CZ2=rdd,options=o); print("smu mocks cross ok",dtype.__name__)
print("done")
