import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
from nebulasem_b200 import host
s = host.Solver.synthetic("bubble3d:1,1,1,0.0003", 100, 100, 100, 4); s.attach(0)
d0 = s.diagnostics()
for blk in range(6):
    s.step(25)
    d = s.diagnostics()
    print('steps', 25*(blk+1), 'courant_max %.4e mass rel change %.2e' % (d['courant_max'], (d['mass']-d0['mass'])/d0['mass']), flush=True)
s.close()
