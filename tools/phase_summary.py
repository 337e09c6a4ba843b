#!/usr/bin/env python
"""Summarise a per-phase cycle profile written by the diagnostic build (-DDI_PROFILE_PHASES, DI_B200_PROF=file):
share of every phase over the launch, and the first / last tiles in Mcycles (CTA cycles summed over work items)."""
import csv
import sys

COLS = ['lookup', 'dense', 'sparse', 'wait_presel', 'scan_emit', 'cut']
for path in sys.argv[1:]:
    rows = list(csv.DictReader(open(path)))
    tot = {c: sum(int(r[c]) for r in rows) for c in COLS}
    total = sum(tot.values())
    print(path, 'tiles', len(rows), 'total Gcycles %.2f' % (total / 1e9), {c: round(100 * tot[c] / total, 1) for c in COLS})
    for r in rows[:8] + rows[-2:]:
        print('  tile', r['tile'], 'items', r['items'], 'Mcycles', round(sum(int(r[c]) for c in COLS) / 1e6, 1),
              [round(int(r[c]) / 1e6, 1) for c in COLS])
