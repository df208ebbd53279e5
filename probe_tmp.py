import sys, os
sys.path.insert(0, '/root/repo')
from nebulasem_b200 import host
s = host.Solver.synthetic("bubble3d", 100, 100, 100, 4); s.attach(0)
s.step(3); s.sync()
ms, pk = s.time_steps(4, per_kernel=True)
print('probe', os.environ.get('NSEM_PROBE'), 'A %.2f ms  B %.2f ms' % (pk[0]/4, pk[2]/4))
