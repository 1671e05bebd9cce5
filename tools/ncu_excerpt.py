"""Excerpt of an .ncu-rep (`ncu --set full` capture of one launch): the metrics DESIGN.md / profiles/README.md cite.
    python tools/ncu_excerpt.py capture.ncu-rep > profiles/rN_full_<kernel>.txt"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(r'^(Kernel Name|Grid Size|Block Size|dram__bytes_(read|write)\.sum($|\.per_second|\.pct)|gpu__time_duration\.sum|'
                  r'gpu__dram_throughput\.avg\.pct|sm__throughput\.avg\.pct|sm__inst_executed_pipe_(tc|tmem|uniform|alu|fma|lsu)[a-z_]*\.(sum|avg)($|\.pct)|'
                  r'sm__pipe_tc[a-z_]*cycles_active\.avg\.pct|sm__pipe_tensor[a-z_0-9]*\.avg\.pct|sm__warps_active\.avg\.pct|'
                  r'launch__(registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit|waves)|'
                  r'l1tex__throughput\.avg\.pct|lts__throughput\.avg\.pct|lts__t_sector_hit_rate\.pct|lts__t_bytes\.sum($|\.per_second)|'
                  r'l1tex__data_bank_conflicts_pipe_lsu\.sum|smsp__cycles_active\.avg|sm__cycles_elapsed\.(avg|max)$|smsp__inst_executed\.sum$|'
                  r'smsp__average_warp[a-z_]*issue_stalled_[a-z_]+_per_warp_active\.pct)')


def main():
    out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    for h, u, v in zip(hdr, units, vals):
        if KEEP.match(h):
            print('%-90s %s %s' % (h, v, u))


if __name__ == '__main__':
    main()
