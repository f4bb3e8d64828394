"""Runs tests/kernel_checks.py CHECKS on the GPU, isolating faults: checks run sequentially in a worker process; if the
worker dies, hangs or hits a CUDA error, a fresh worker continues after the offending check.
    python tools/run_checks.py [--only substr[,substr]] [--knobs 2=1,1=64] [--out gpurun_out/checks.json]
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WORKER = r'''
import json, os, sys, time, traceback
sys.path.insert(0, %(root)r)
import torch
from contrastive_masked_unet_b200._lib import lib
from tests.kernel_checks import CHECKS
for kv in filter(None, os.environ.get('CMU_KNOBS', '').split(',')):
    k, v = kv.split('='); lib.cmu_debug_set(int(k), int(v))
names = sys.argv[1:]
for n in names:
    t0 = time.time()
    try:
        res = CHECKS[n]()
        torch.cuda.synchronize()
        print('RESULT ' + json.dumps({'name': n, 'ok': True, 'res': res, 's': round(time.time() - t0, 2)}), flush=True)
    except AssertionError as e:
        print('RESULT ' + json.dumps({'name': n, 'ok': False, 'res': str(e)[:600], 's': round(time.time() - t0, 2)}), flush=True)
        try:
            torch.cuda.synchronize()
        except Exception as e2:
            print('RESULT ' + json.dumps({'name': n + ':post', 'ok': False, 'fatal': True, 'res': str(e2)[:300]}), flush=True)
            sys.exit(3)
    except Exception as e:
        print('RESULT ' + json.dumps({'name': n, 'ok': False, 'fatal': True, 'res': (type(e).__name__ + ': ' + str(e))[:600]}), flush=True)
        sys.exit(3)
'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default='')
    ap.add_argument('--knobs', default='')
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'checks.json'))
    ap.add_argument('--timeout', type=int, default=150)
    args = ap.parse_args()
    from tests.kernel_checks import CHECKS  # noqa: names only (imports torch)
    names = list(CHECKS)
    if args.only:
        subs = args.only.split(',')
        names = [n for n in names if any(s in n for s in subs)]
    env = dict(os.environ, CMU_KNOBS=args.knobs)
    results = []
    todo = names
    while todo:
        p = subprocess.Popen([sys.executable, '-c', WORKER % {'root': ROOT}] + todo, stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True, env=env)
        done = []
        t_last = time.time()
        tail = []
        import selectors
        sel = selectors.DefaultSelector()
        sel.register(p.stdout, selectors.EVENT_READ)
        hung = False
        while True:
            ev = sel.select(timeout=5)
            if ev:
                line = p.stdout.readline()
                if not line:
                    break
                if line.startswith('RESULT '):
                    r = json.loads(line[7:])
                    results.append(r)
                    done.append(r['name'].split(':')[0])
                    print(('PASS ' if r['ok'] else 'FAIL ') + r['name'], json.dumps(r['res'])[:700], flush=True)
                    t_last = time.time()
                else:
                    tail.append(line.rstrip())
                    tail = tail[-15:]
            elif time.time() - t_last > args.timeout + (60 if not done else 0):
                hung = True
                p.kill()
                break
        p.wait()
        remaining = [n for n in todo if n not in done]
        last_fatal = bool(results) and results[-1].get('fatal', False) and results[-1]['name'].split(':')[0] in done
        if hung or (p.returncode != 0 and not last_fatal):
            if remaining:
                bad = remaining[0]
                results.append({'name': bad, 'ok': False, 'fatal': True,
                                'res': ('HUNG' if hung else f'worker exit {p.returncode}') + ' | ' + ' / '.join(tail[-6:])})
                print('FAIL ' + bad, results[-1]['res'][:900], flush=True)
                todo = remaining[1:]
            else:
                todo = []
        elif p.returncode != 0:
            todo = remaining
        else:
            todo = []
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({'knobs': args.knobs, 'results': results}, open(args.out, 'w'), indent=1)
    n_ok = sum(1 for r in results if r['ok'])
    print(f'SUMMARY knobs={args.knobs!r}: {n_ok}/{len(results)} passed')
    return 0 if n_ok == len(results) else 1


if __name__ == '__main__':
    sys.exit(main())
