import sys, importlib, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import numpy as np
pkg = importlib.import_module("orb_slam2_modification_with-point-and-line-feature_b200")
api = pkg.load_api()
img = pkg.synth.frame(1000,640,480)
ex = api.LineExtractor()
ex.ExtractLineSegment(img)
lines,width,prec,nfa = ex.lsd_segments()
g = np.load('/root/repo/tests/golden/lsd_cfgA_seed1000.npz')
d = np.abs(nfa-g['nfa'])
bad = np.where(d>1e-6)[0]
print('n bad', len(bad), 'of', len(nfa))
for i in bad[:15]:
    L = np.hypot(*(lines[i,:2]-lines[i,2:]))
    print(i, 'gpu', nfa[i], 'ref', g['nfa'][i], 'p', prec[i], 'w', width[i], 'len', L, 'line', lines[i])
