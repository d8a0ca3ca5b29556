#!/usr/bin/env python
"""Decode the S3D_TRACE event timeline printed by an experiment build of the decoder kernel (see decoder_tc.cu)."""
import sys
names = {10: 'I  h_ready seen', 11: 'I  slot A2 full', 12: 'I  A2 issued', 13: 'I  slot B2 full', 14: 'I  B2 issued + h_free commit',
         16: 'I  slot A1 full', 17: 'I  A1 issued', 19: 'I  slot B1 full', 20: 'I  B1 issued + d1_ready commit',
         30: 'C  d1_ready seen', 31: 'C  math done', 32: 'C  h_free seen', 33: 'C  H stored', 34: 'C  arrived'}
ev = []
for l in open(sys.argv[1]):
    if l.startswith('TR'):
        _, role, e, c, t = l.split()
        ev.append((int(t), int(role), int(e), int(c)))
if not ev:
    sys.exit("no events")
# 22-bit clock: unwrap per role in program order
out = []
for role in (0, 1):
    last, base = None, 0
    for t, r, e, c in [x for x in ev if x[1] == role]:
        if last is not None and t < last - (1 << 21):
            base += 1 << 22
        last = t
        out.append((t + base, r, e, c))
out.sort()
t0 = out[0][0]
prev = {0: None, 1: None}
for t, r, e, c in out:
    d = t - prev[r] if prev[r] is not None else 0
    print(f"{t - t0:7d} {'' if r == 0 else ' ' * 40}{names.get(e, e)} c={c} (+{d})")
    prev[r] = t
