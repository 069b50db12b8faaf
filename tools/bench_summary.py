"""Prints the headline fields of a bench.py JSON line (file argument)."""
import json
import sys

d = json.load(open(sys.argv[1]))
r, t = d['roofline'], d['tree_roofline']
print('value %.4g  e2e %.4g  ms/step %.2f  evals/s %.4g  rows/sim %.3f' % (d['value'], (d['e2e'] or {'value': float('nan')})['value'], d['ms_per_step'],
      d['evals_per_second'], d['sims_breakdown']['network_rows']))
print('tower ms %.4f rows %.0f frac %.3f mma_frac %.3f | tree ms %.4f | nocache %.4g | clocks %s' % (
    r['ms_per_launch'], r['rows_per_launch'], r['frac'], r['mma_frac'], t['ms_per_launch'], (d['without_cache_lockstep'] or {'value': float('nan')})['value'], d['clocks']))
