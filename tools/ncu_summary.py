"""Selected counters of every kernel in .ncu-rep files -> JSON (profiles/*.json).

    python tools/ncu_summary.py out.json report1.ncu-rep [report2.ncu-rep ...]
"""
import csv
import json
import subprocess
import sys

KEYS = [
  'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
  'sm__inst_executed.avg.per_cycle_active', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
  'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
  'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
  'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
  'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
  'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
  'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
  'sm__throughput.avg.pct_of_peak_sustained_elapsed',
  'dram__throughput.avg.pct_of_peak_sustained_elapsed',
  'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
  'sm__cycles_elapsed.avg.per_second',
]


def main():
  out = {}
  for rep in sys.argv[2:]:
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    head, units = rows[0], rows[1]
    for row in rows[2:]:
      d = dict(zip(head, row))
      u = dict(zip(head, units))
      name = d.get('Kernel Name', '?')
      entry = {'report': rep.split('/')[-1]}
      for k in KEYS:
        if k in d:
          entry[k] = d[k] + (' ' + u[k] if u.get(k) else '')
      out[name if name not in out else name + ' [' + entry['report'] + ']'] = entry
  with open(sys.argv[1], 'w') as f:
    json.dump(out, f, indent=1)
  print(json.dumps(out, indent=1))


if __name__ == '__main__':
  main()
