#!/usr/bin/env python
"""Run every `-m gpu` test FUNCTION in its own process with a timeout, so one trapping kernel (dead CUDA context)
or hang cannot take the rest of the suite with it.  Logs go to gpurun_out/tests/.  Development tool for gpurun calls:
    gpurun -- 'python tools/gpu_tests_isolated.py [-k substr] [--timeout 240]'
"""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('-k', default='')
    ap.add_argument('--timeout', type=int, default=240)
    ap.add_argument('--files', nargs='*', default=['tests'])
    a = ap.parse_args()
    out_dir = os.path.join(ROOT, 'gpurun_out', 'tests')
    os.makedirs(out_dir, exist_ok=True)
    r = subprocess.run([sys.executable, '-m', 'pytest', '--collect-only', '-q', '-m', 'gpu'] + a.files, cwd=ROOT,
                       capture_output=True, text=True)
    funcs = []
    for line in r.stdout.splitlines():
        if '::' in line:
            f = line.split('[')[0].strip()
            if f not in funcs and a.k in f:
                funcs.append(f)
    print(f'{len(funcs)} test functions', flush=True)
    summary = []
    for f in funcs:
        name = f.replace('/', '_').replace('::', '-')
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, '-m', 'pytest', f, '-q', '--no-header', '-p', 'no:cacheprovider', '-m', 'gpu', '-rP'],
                               cwd=ROOT, capture_output=True, text=True, timeout=a.timeout)
            rc, out = p.returncode, p.stdout + p.stderr
        except subprocess.TimeoutExpired as e:
            rc, out = -9, (e.stdout or b'').decode(errors='replace') + (e.stderr or b'').decode(errors='replace') + '\nTIMEOUT'
        with open(os.path.join(out_dir, name + '.log'), 'w') as fh:
            fh.write(out)
        tail = [l for l in out.strip().splitlines() if l.strip()][-1:] or ['']
        line = f'{"PASS" if rc == 0 else "FAIL"} rc={rc} {time.time() - t0:6.1f}s {f} :: {tail[0][:160]}'
        print(line, flush=True)
        summary.append(line)
    with open(os.path.join(out_dir, 'SUMMARY.txt'), 'w') as fh:
        fh.write('\n'.join(summary) + '\n')
    sys.exit(0 if all(s.startswith('PASS') for s in summary) else 1)


if __name__ == '__main__':
    main()
