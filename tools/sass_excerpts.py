"""SASS excerpts for profiles/ from the built objects (cuobjdump -sass; no GPU needed):
   python tools/sass_excerpts.py  ->  profiles/sass_*_r02.txt
   * the 8-term lock-step loop of the heavy ridge kernel (FP64 / other instruction mix per term)
   * the shared-memory histogram update (ATOMS.POPC.INC + the LDS / DADD / ATOMS.CAST.SPIN loop) of the light kernel
   * the bulk-copy (UBLKCP) / mbarrier (SYNCS) instructions of k_reduce"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = os.path.join(ROOT, 'vegas_b200', 'csrc')
OUT = os.path.join(ROOT, 'profiles')


def sass(obj, function):
    txt = subprocess.run(['cuobjdump', '-sass', os.path.join(CS, obj)], capture_output=True, text=True).stdout
    lines, on = [], False
    for l in txt.splitlines():
        if 'Function :' in l:
            on = function in l
            continue
        if on:
            m = re.match(r'^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*/\*', l)
            if m:
                lines.append('%s  %s' % (m.group(1), m.group(2).rstrip(' ;')))
    return lines


def is_fp64(l):
    return re.search(r'\b(DFMA|DADD|DMUL)\b', l) is not None


def _unused_densest(lines, width):
    d = [1 if is_fp64(l) else 0 for l in lines]
    s, best = sum(d[:width]), (0, 0)
    for i in range(len(lines) - width):
        if s > best[0]:
            best = (s, i)
        s += d[i + width] - d[i]
    return best[1]


def main():
    # ---- heavy ridge, exact 8-D instantiation
    L = sass('fused_ridge.o', '_Z8k_engineI8FusedSrcI6FRidgeLi8ELb0ELb0ELb1EEEv7EnginePT_')
    # the loop ends with the one branch on the out-of-range check of all 8 arguments (BRA P2: two predicates);
    # it starts at the BSSY ~270 instructions before (the loop of the parameter-bank centres comes first in the code)
    cand = [n for n, l in enumerate(L) if re.search(r'BRA P\d, 0x', l) and n > 260]
    j = max(cand, key=lambda n: sum(is_fp64(l) for l in L[n - 260:n]))
    i = j - 200
    while i > 0 and 'BSSY' not in L[i]:
        i -= 1
    body = L[i:j + 1]
    nf = sum(is_fp64(l) for l in body)
    with open(os.path.join(OUT, 'sass_ridge_term_loop_r02.txt'), 'w') as fh:
        fh.write('# k_engine<FusedSrc<FRidge, 8, heavy, exact>>: one pass of the 8-wide lock-step term loop\n'
                 '# (integrands.cuh: sum_axis_order_par + common.cuh: vb_exp_n<8>), cuobjdump -sass of fused_ridge.o\n'
                 '# %d instructions: %d FP64 (DFMA/DADD/DMUL) + %d other = %.1f FP64 and %.1f other per term\n'
                 % (len(body), nf, len(body) - nf, nf / 8., (len(body) - nf) / 8.))
        fh.write('\n'.join(body) + '\n')
    # ---- light ridge: histogram update
    L = sass('fused_light.o', '_Z8k_engineI8FusedSrcI11FRidgeLightLi8ELb1ELb1ELb1EEEv7EnginePT_')
    k = next(n for n, l in enumerate(L) if 'ATOMS.CAST.SPIN' in l)
    with open(os.path.join(OUT, 'sass_hist_update_r02.txt'), 'w') as fh:
        fh.write('# k_engine<FusedSrc<FRidgeLight, 8, light, exact>>: training-histogram update of one axis (engine.cuh:\n'
                 '# hist_add_code): native u32 count (ATOMS.POPC.INC) + fp64 sum as LDS.64 / DADD / ATOMS.CAST.SPIN.64 / BRA\n')
        fh.write('\n'.join(L[k - 14:k + 4]) + '\n')
    # ---- k_reduce: bulk copies and mbarriers
    L = sass('reduce_buffer.o', '_Z8k_reduceILi1EEv7EngineP')
    with open(os.path.join(OUT, 'sass_reduce_bulk_copy_r02.txt'), 'w') as fh:
        fh.write('# k_reduce<1>: the TMA 1-D bulk copies (UBLKCP) into the shared-memory stages and their mbarriers (SYNCS),\n'
                 '# cuobjdump -sass of reduce_buffer.o (reduce.cuh: issue / arrive)\n')
        for n, l in enumerate(L):
            if re.search(r'UBLKCP|SYNCS|FENCE.VIEW.ASYNC', l):
                fh.write(l + '\n')
    print('written: sass_ridge_term_loop_r02.txt sass_hist_update_r02.txt sass_reduce_bulk_copy_r02.txt')


if __name__ == '__main__':
    main()
