"""Generates the kernels of the instruction-supply probe (tools/ifetch_probe.cu): `probe` executes N
straight-line FP64 statements (8 independent DMUL/DADD chains, no loop, no memory traffic), cut into
__noinline__ functions of 2048 statements so ptxas stays linear in N.

usage: gen_ifetch_probe.py N > rN.cu ; nvcc -cubin -arch=sm_100a --fmad=false -o tools/ifetch/rN.cubin rN.cu
       (N = 32768, 131072, 524288 -> 0.5, 2, 8 MB of code; results: profiles/r01_ifetch_probe.jsonl)"""
import sys

n = int(sys.argv[1])
per = 2048
print("struct V8 { double v[8]; };")
for f in range(n // per):
    print("static __device__ __noinline__ void f%d(V8* s, double m, double a) {" % f)
    print("  double v0=s->v[0], v1=s->v[1], v2=s->v[2], v3=s->v[3], v4=s->v[4], v5=s->v[5], v6=s->v[6], v7=s->v[7];")
    for _ in range(per // 8):
        print("  v0=__dmul_rn(v0,m); v1=__dadd_rn(v1,a); v2=__dmul_rn(v2,m); v3=__dadd_rn(v3,a); "
              "v4=__dmul_rn(v4,m); v5=__dadd_rn(v5,a); v6=__dmul_rn(v6,m); v7=__dadd_rn(v7,a);")
    print("  s->v[0]=v0; s->v[1]=v1; s->v[2]=v2; s->v[3]=v3; s->v[4]=v4; s->v[5]=v5; s->v[6]=v6; s->v[7]=v7;\n}")
print('extern "C" __global__ void __launch_bounds__(256,2) probe(double* sink, double m, double a) {')
print("  V8 s; for (int k=0;k<8;k++) s.v[k]=threadIdx.x*1e-9+k;")
for f in range(n // per):
    print("  f%d(&s, m, a);" % f)
print("  double t=0; for (int k=0;k<8;k++) t+=s.v[k]; if (t==123.456) sink[0]=t;")
print("}")
