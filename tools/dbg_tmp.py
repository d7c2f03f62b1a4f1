import sys; sys.path.insert(0,'/root/repo')
import rakau_b200 as rk
m,x,y,z=rk.plummer(4000000)
g=rk.Octree(); g.build(x,y,z,m); g.acc_pot(0,0.75); print(g.eval_info.asdict())
