"""per-source-line view of an .ncu-rep (needs -lineinfo and --import-source on):
   python tools/ncu_lines.py prof.ncu-rep [top]  -> lines sorted by stall samples, with executed instructions"""
import csv
import subprocess
import sys


def main(path, top=40):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname, hdr, lines = '', None, []
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path':
            fname = r[1].split('/')[-1]
        elif len(r) > 2 and r[0] == 'Line No':
            hdr = r
        elif hdr and len(r) > 8 and r[0] != '':
            try:
                lines.append((int(r[4]), int(r[7]), fname, int(r[0]), r[1].strip()))
            except ValueError:
                pass
    tot_s = sum(l[0] for l in lines) or 1
    tot_i = sum(l[1] for l in lines) or 1
    print('total samples %d  total warp instructions %d' % (tot_s, tot_i))
    for s, i, f, n, src in sorted(lines, reverse=True)[:top]:
        print('%5.1f%% smp %5.1f%% inst  %-14s:%-4d %s' % (100. * s / tot_s, 100. * i / tot_i, f, n, src[:110]))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
