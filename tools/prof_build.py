"""Profile target: 3 builds of the 4M Plummer tree (used under ncu for per-kernel build times)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rakau_b200 as rk
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
m, x, y, z = rk.plummer(N)
g = rk.Octree()
for _ in range(3):
    bi = g.build(x, y, z, m)
print(bi.asdict())
