"""Profile target: one 4M build + 3 traversals (used under ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rakau_b200 as rk
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 128
m, x, y, z = rk.plummer(N)
g = rk.Octree(); g.build(x, y, z, m, ncrit=nc)
for _ in range(3):
    g.acc_pot(0, 0.75)
print(g.eval_info.asdict())
