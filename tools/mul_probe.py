import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_groth16_prover_3x_b200 as pkg
curves = (0, 1) if len(sys.argv) < 2 else (int(sys.argv[1]),)
for curve in curves:
    ctx = pkg.MsmContext(curve, 0)
    print("curve", curve, "fq modmul (registers, 32 warps/SM) G/s %.3f" % ctx.microbench(2, 1024))
    for kind in (7, 8, 9, 11, 12, 13):
        print("  slab mul: group %d blocks/SM %d  %.3f G tower-mul/s" % (1 if kind < 11 else 2, (kind - 7) % 4 + 1, ctx.microbench(kind, 512)))
    ctx.close()
