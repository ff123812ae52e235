"""derive the degree-11 near-minimax polynomial for exp(r), |r| <= ln2/2, used by vb_exp_n
(vegas_b200/csrc/common.cuh)"""
import mpmath as mp
mp.mp.dps = 60
a = mp.log(2) / 2 * mp.mpf('1.0001')
c = mp.chebyfit(mp.exp, [-a, a], 12)
err = max(abs(mp.polyval(c, -a + 2 * a * i / 2000) / mp.exp(-a + 2 * a * i / 2000) - 1) for i in range(2001))
print('max rel err', mp.nstr(err, 5))
for i, ci in enumerate(c):
    print('%.17e  r^%d' % (float(ci), 11 - i))
