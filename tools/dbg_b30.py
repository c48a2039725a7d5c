import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from helpers import *
from uisrnn_b200 import native
m = native.NativeModel(load_weights('model_small.npz'))
case = [c for c in small_cases() if c['name']=='b30'][0]
labs, dbg = m.predict([case['x']], beam_size=30, test_iteration=2, trace_utt=0)
w, gw = dbg['win'], case['win']
bad = np.where((w != gw).any(axis=1))[0]
print('n rows', len(w), 'n bad', len(bad), 'first bad', bad[:10])
off = case['off']
r0 = bad[0]; step = np.searchsorted(off, r0, side='right')-1
print('step', step, 'rows', off[step], off[step+1])
sl = slice(off[step], off[step+1])
for r in range(off[step], off[step+1]):
    print(r-off[step], w[r], gw[r], '%.6f %.6f' % (dbg['score'][r], case['score'][r]))
