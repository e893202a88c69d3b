import sys; sys.path.insert(0,'/root/repo')
import gpu_groth16_prover_3x_b200 as pkg
ctx = pkg.MsmContext(0, 0)
for k in (4,5,6):
    print("kind", k, "us per inversion", ctx.microbench(k, 50))
