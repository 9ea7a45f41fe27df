import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dizk_oracle as O
from oracle import ref_cuda as R
log = open("gpurun_out/diag.log", "a", buffering=1)
def p(*a):
    print(*a, file=log, flush=True)
p("start")
sc = [1, 2, 3, 4]
t = time.time(); out = R.field_batch(O.pack_scalars(sc) + O.le32(7), 4); p("field_batch ok", time.time() - t, [int.from_bytes(out[64*i:64*i+64], "big") for i in range(4)])
base = O.G1.random(10)
for (ss, w, n) in ((253, 5, 4), (253, 13, 64)):
    outerc = (ss + w - 1) // w
    t = time.time()
    out = R.fixed_batch(outerc, w, outerc, 1 << w, n, ss, O.pack_g1([base]), O.pack_scalars(list(range(1, n + 1))), 1)
    pts = O.unpack_g1(out, stride=64, big_endian=True)
    p("fixed_batch", ss, w, n, time.time() - t, O.G1.equals(pts[2], O.G1.mul(base, 3)))
p("done")
