"""Turn the scratch artefacts of a tools/gpu_round.sh call (gpurun_out/) into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r01        # -> profiles/r01_launches.csv, r01_launches_summary.md, r01_ncu_full_summary.md,
                                                  #    r01_bench_n1.json, r01_bench_reference.json, ncu_traffic.json
Reads .ncu-rep files with `ncu -i ... --page raw --csv` (ncu is in the image; no GPU needed).
"""
import collections
import csv
import glob
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')

KEY = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
       'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max',
       'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
       'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
       'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
       'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
       'sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
       'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
       'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
       'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
       'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
       'smsp__sass_inst_executed_op_utcmma.sum', 'smsp__sass_inst_executed_op_tmem_ldt.sum']


def launches(tag):
    src = os.path.join(OUT, 'launches.csv')
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(PROF, tag + '_launches.csv'))
    lines = [l for l in open(src) if not l.startswith('==')]
    agg, tot, n = collections.OrderedDict(), 0.0, 0
    for row in csv.DictReader(io.StringIO(''.join(lines))):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v = v / 1000 if row['Metric Unit'] == 'ns' else v * 1000 if row['Metric Unit'] == 'ms' else v
        a = agg.setdefault(row['Kernel Name'][:72], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
    with open(os.path.join(PROF, tag + '_launches_summary.md'), 'w') as f:
        f.write('# %s launch list: `ncu --metrics gpu__time_duration.sum --clock-control none -c 700 python bench.py --steps 2 --warmup 1 '
                '--no-graph ...` (tools/gpu_round.sh)\n# cold-cache, serialised launches: compare SHARES, not absolutes; includes the set-up '
                'launches of the bench\n\n| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n' % tag)
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
            f.write('| `%s` | %d | %.1f | %.1f%% | %.1f |\n' % (k, c, t, 100 * t / tot, t / c))
        f.write('\ntotal %.1f us over %d launches\n' % (tot, n))


def stage_launches(tag):
    """gpurun_out/launches_<stage>.csv (tools/gpu_call.sh stages) -> profiles/<tag>_launches_<stage>.md: per-kernel totals of one stage run."""
    for src in sorted(glob.glob(os.path.join(OUT, 'launches_*.csv'))):
        stage = os.path.basename(src)[len('launches_'):-4]
        lines = [l for l in open(src) if not l.startswith('==')]
        agg, tot, n = collections.OrderedDict(), 0.0, 0
        for row in csv.DictReader(io.StringIO(''.join(lines))):
            if row.get('Metric Name') != 'gpu__time_duration.sum':
                continue
            v = float(row['Metric Value'].replace(',', ''))
            v = v / 1000 if row['Metric Unit'] == 'ns' else v * 1000 if row['Metric Unit'] == 'ms' else v
            a = agg.setdefault(row['Kernel Name'].split('(')[0][:72], [0, 0.0])
            a[0] += 1
            a[1] += v
            tot += v
            n += 1
        if not n:
            continue
        with open(os.path.join(PROF, '%s_launches_%s_final.md' % (tag, stage)), 'w') as f:
            f.write('# %s launch list of `python tools/run_stage.py %s` under `ncu --metrics gpu__time_duration.sum --clock-control none`\n'
                    '# cold-cache, serialised launches, set-up launches of the stage included: compare SHARES, not absolutes\n\n'
                    '| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n' % (tag, stage))
            for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
                f.write('| `%s` | %d | %.1f | %.1f%% | %.1f |\n' % (k, c, t, 100 * t / tot, t / c))
            f.write('\ntotal %.1f us over %d launches\n' % (tot, n))


def full(tag):
    traffic = {}
    path = os.path.join(PROF, 'ncu_traffic.json')
    if os.path.exists(path):
        traffic = json.load(open(path))
    out = ['# %s ncu --set full captures (clock-control none).  Raw .ncu-rep files are in gpurun_out/ (scratch); key metrics below.\n' % tag]
    for rep in sorted(glob.glob(os.path.join(OUT, '*.ncu-rep'))):
        r = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(r.stdout)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for row in rows[2:]:
            name = row[hdr.index('Kernel Name')]
            out.append('\n## %s  (%s)\n' % (name.split('(')[0][:80], os.path.basename(rep)))
            vals = {}
            for k in KEY:
                if k in hdr:
                    vals[k] = row[hdr.index(k)]
                    out.append('- `%s` = %s %s' % (k, row[hdr.index(k)], units[hdr.index(k)]))
            try:
                scale = {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1.0}
                rd = float(vals['dram__bytes_read.sum']) * scale[units[hdr.index('dram__bytes_read.sum')]]
                wr = float(vals['dram__bytes_write.sum']) * scale[units[hdr.index('dram__bytes_write.sum')]]
                out.append('- dram traffic (read+write) = %.1f MB' % ((rd + wr) / 1e6))
                short = name.split('<')[0].split('(')[0].replace('void ', '').replace('lemo::', '')
                traffic[short] = rd + wr
            except Exception:
                pass
    # one lemo_smplx_forward = pose/chain + blend GEMM + skinning GEMM + output joints (bench.py: roofline_lbs.traffic)
    parts = ['k_pose_chain_fwd', 'k_blend_v2', 'k_skin_tc', 'k_joints_fwd']
    if all(p in traffic for p in parts):
        traffic['lbs_forward'] = sum(traffic[p] for p in parts)
    traffic['_source'] = 'profiles/%s_ncu_full_summary.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)' % tag
    open(os.path.join(PROF, tag + '_ncu_full_summary.md'), 'w').write('\n'.join(out) + '\n')
    json.dump(traffic, open(path, 'w'), indent=1)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    stage_launches(tag)
    full(tag)
    for src, dst in (('bench_n1.json', '_bench_n1.json'), ('bench_ref.json', '_bench_reference.json'), ('bench_n2.json', '_bench_n2.json'),
                     ('sanitizer_memcheck.log', '_sanitizer_memcheck.log'), ('sanitizer_racecheck_perframe.log', '_sanitizer_racecheck_perframe.log'),
                     ('sanitizer_racecheck_prox.log', '_sanitizer_racecheck_prox.log'), ('sanitizer_racecheck_infill.log', '_sanitizer_racecheck_infill.log'),
                     ('smoke.log', '_smoke.log'), ('diag_lbs.log', '_lbs_forward_timing.txt')):
        p = os.path.join(OUT, src)
        if os.path.exists(p) and os.path.getsize(p) > 0:
            shutil.copy(p, os.path.join(PROF, tag + dst))


if __name__ == '__main__':
    main()
