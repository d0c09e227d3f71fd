"""Development aid: digest of the chain kernel's event log (scripts/chain_trace.py): per layer of one steady-state chain, when its MMAs
were issued and its epilogues ran, how long the issuers waited for weights / accumulators / operands."""
import sys
lines = open(sys.argv[1]).read().split('\n')
idx = [i for i, l in enumerate(lines) if l.startswith('EVAL')]
seg = lines[idx[0] + 1:idx[1]]
ev = []
for l in seg:
    if l.startswith('T '):
        _, r, t, c = l.split(); ev.append((int(c), int(r), int(t, 16)))
ev.sort()
starts = [c for c, r, t in ev if r == 0 and t == 0x100]
print('chain durations:', [b - a for a, b in zip(starts, starts[1:])][:14])
which = int(sys.argv[2]) if len(sys.argv) > 2 else 3
a, b = starts[which], starts[which + 1]
cur = [e for e in ev if a <= e[0] < b + 3000]
# per issuer: time waiting on bFull (0xa00 -> 0x400), on accEmpty (0x100 -> 0x200)
for g in (0, 1):
    wb = wa = 0; last = {}
    for c, r, t in cur:
        if r != g: continue
        k = t >> 8
        if k == 0xa: last['b'] = c
        if k == 4 and 'b' in last: wb += c - last.pop('b')
        if k == 1: last['a'] = c
        if k == 2 and 'a' in last: wa += c - last.pop('a')
    print('issuer %d: waiting for weights %d clk, for an accumulator stage %d clk' % (g, wb, wa))
for j in range(6):
    iss = [c - a for c, r, t in cur if r < 2 and (t >> 8) in (2, 4, 5) and ((t >> 4) & 15) == j and c < b + 3000]
    epi = [(c - a, t >> 8) for c, r, t in cur if r >= 2 and (t >> 8) in (7, 8) and ((t >> 4) & 15) == j]
    opr = [c - a for c, r, t in cur if r >= 2 and (t >> 8) == 9 and ((t >> 4) & 15) == j]
    print('layer %d: issue %6d .. %6d   epilogues got %s  released %s  published %s' % (j, min(iss), max(iss), [c for c, k in epi if k == 7], [c for c, k in epi if k == 8], opr))
if len(sys.argv) > 3:
    names = {1: 'I.waitAccEmpty', 2: 'I.gotAcc', 3: 'I.opReady', 4: 'I.bFull', 5: 'I.commitAcc', 6: 'E.waitFull', 7: 'E.gotFull', 8: 'E.accEmptyArr', 9: 'E.opReadyArr', 10: 'I.waitB'}
    for c, r, t in cur:
        print('%7d  r%d  %-16s j=%d c/kc=%d' % (c - a, r, names[t >> 8], (t >> 4) & 15, t & 15))
