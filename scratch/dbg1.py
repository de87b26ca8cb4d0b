import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/rt-gaussian-splat-renderer_b200'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from oracle import lbvh_ref as L
from gpu_util import random_set, make_scene
for n in (1,2,3,16,255,2049,5000,100000):
    gs=random_set(n, 100+n, sh=False)
    sc=make_scene(gs); lb=sc.read_lbvh(); ref=L.build(gs.pos)
    print(n, 'morton', np.array_equal(lb['morton'],ref['codes']), 'sorted', np.array_equal(lb['sorted_idx'],ref['sorted_idx']),
          'child', np.array_equal(lb['child'],ref['child']), 'parent', np.array_equal(lb['parent'],ref['parent']))
    if not np.array_equal(lb['morton'],ref['codes']):
        bad=np.nonzero(lb['morton']!=ref['codes'])[0]; print('  nbad',len(bad), bad[:5], lb['morton'][bad[:5]], ref['codes'][bad[:5]])
    if not np.array_equal(lb['sorted_idx'],ref['sorted_idx']):
        bad=np.nonzero(lb['sorted_idx']!=ref['sorted_idx'])[0]; print('  sorted nbad',len(bad), bad[:8], lb['sorted_idx'][bad[:8]], ref['sorted_idx'][bad[:8]])
        print('  is perm', np.array_equal(np.sort(lb['sorted_idx']), np.arange(n)))
        k=ref['codes'][lb['sorted_idx']]; print('  codes nondecreasing', (np.diff(k.astype(np.int64))>=0).all())
