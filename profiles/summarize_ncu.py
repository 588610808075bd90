"""Condense an `ncu --set full` report into the JSON summaries kept under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_ncu_full_v3_summary.json [n_envs]

Per captured launch: duration, warp instructions (total and per env), lane utilisation, issue utilisation, achieved
occupancy, registers / shared memory, DRAM bytes read + written (the `traffic` of bench.py's roofline block), L1 / L2 hit
rates and the warp-stall breakdown (share of sampled stall reasons)."""
import csv
import io
import json
import re
import subprocess
import sys

STAGES = [('kpos_p0', 'pos'), ('kcol_', 'col'), ('kcon_p0', 'proj'), ('kvel_p0', 'vel'), ('kact_p0', 'smooth'),
          ('fb_run_solve', 'solve'), ('kfin_f1', 'finish'), ('ph_scatter', 'misc'), ('ph_pack', 'pack'), ('ph_reset', 'misc')]


def stage_of(name):
    for k, v in STAGES:
        if k in name:
            return v
    return name[:40]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    n_envs = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}

    def f(r, name, default=None):
        try:
            return float(r[col[name]].replace(',', ''))
        except Exception:
            return default
    to_bytes = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    to_us = {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'msecond': 1e3, 'usecond': 1, 'nsecond': 1e-3, 'second': 1e6}
    res = []
    for r in data:
        name = r[col['Kernel Name']]
        dur = f(r, 'gpu__time_duration.sum') * to_us[units[col['gpu__time_duration.sum']]]
        rd = f(r, 'dram__bytes_read.sum') * to_bytes[units[col['dram__bytes_read.sum']]]
        wr = f(r, 'dram__bytes_write.sum') * to_bytes[units[col['dram__bytes_write.sum']]]
        inst = f(r, 'smsp__inst_executed.sum')
        stalls = {}
        for n, i in col.items():
            m = re.match(r'smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio', n)
            if m and 'not_issued' not in n:
                stalls[m.group(1)] = float(r[i] or 0)
        tot = sum(stalls.values()) or 1.0
        top = {k: round(v / tot, 3) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:5]}
        res.append({
            'stage': stage_of(name), 'kernel': name[:120], 'duration_us': round(dur, 2),
            'grid': r[col['launch__grid_size']].strip(), 'block': r[col['launch__block_size']].strip(),
            'registers_per_thread': int(f(r, 'launch__registers_per_thread')),
            'dyn_smem_per_block_bytes': f(r, 'launch__shared_mem_per_block_dynamic') * to_bytes.get(units[col['launch__shared_mem_per_block_dynamic']], 1),
            'waves_per_sm': f(r, 'launch__waves_per_multiprocessor'),
            'warp_instructions': int(inst), 'warp_instructions_per_env': round(inst / n_envs, 1),
            'active_threads_per_instruction': f(r, 'smsp__thread_inst_executed_per_inst_executed.ratio'),
            'issue_active_pct': f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
            'achieved_warps_pct_of_64': f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'),
            'sm_throughput_pct': f(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'),
            'dram_bytes_read': int(rd), 'dram_bytes_write': int(wr), 'dram_bytes_per_launch': int(rd + wr),
            'dram_bytes_per_env': round((rd + wr) / n_envs, 1),
            'dram_throughput_pct': f(r, 'dram__throughput.avg.pct_of_peak_sustained_elapsed'),
            'dram_gbs': round((rd + wr) / (dur * 1e-6) / 1e9, 1),
            'l1_hit_pct': f(r, 'l1tex__t_sector_hit_rate.pct'), 'l2_hit_pct': f(r, 'lts__t_sector_hit_rate.pct'),
            'stall_share_top5': top,
        })
    json.dump({'report': rep.split('/')[-1], 'n_envs': n_envs, 'launches': res}, open(out, 'w'), indent=1)
    for x in res:
        print(f"{x['stage']:8s} {x['duration_us']:8.1f} us  {x['warp_instructions_per_env']:8.0f} inst/env  lanes {x['active_threads_per_instruction']:5.1f}  "
              f"issue {x['issue_active_pct']:5.1f}%  dram {x['dram_bytes_per_launch'] / 1e6:7.1f} MB ({x['dram_gbs']} GB/s)  {x['stall_share_top5']}")


if __name__ == '__main__':
    main()
