import sys, numpy as np
t = np.array([int(x) for x in open(sys.argv[1]).read().split()], dtype=np.int64)
n_layers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
per = 2 + 4 * (n_layers - 1) + 2
names = ["start->L1 issued(incl. wait for image)", "L1 MMA wait"]
for l in range(1, n_layers):
    names += [f"relunorm{l} (own)", f"relunorm{l} wg barrier", f"issue MMA{l+1}", f"MMA{l+1} wait"]
names += ["final epilogue+store"]
nt = len(t) // per
d = np.diff(t[: nt * per].reshape(nt, per), axis=1)
print(f"{nt} tiles of warpgroup 0 / CTA 0; cycles per phase (median over tiles 2..):")
for k, nme in enumerate(names):
    print(f"  {nme:42s} {np.median(d[2:, k]):8.0f}   (min {d[2:, k].min():6d} max {d[2:, k].max():6d})")
print("  tile period (start to next start):", np.median(np.diff(t[0::per][:nt])))
