"""Turn the ncu artefacts of a round into the tracked summaries under profiles/.

    ncu -i gpurun_out/prof_k1t_<tag>.ncu-rep --page raw --csv > /tmp/k1t_raw.csv
    python tools/summarise_profiles.py <tag> /tmp/k1t_raw.csv gpurun_out/launches_<tag>.csv
"""
import collections, csv, json, sys

tag, raw_csv, launches_csv = sys.argv[1:4]
rows = list(csv.reader(open(raw_csv)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__time_duration.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__block_size', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active']
want += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
out, d = [['metric', 'unit', 'value']], {}
for h in want:
    if h in hdr:
        i = hdr.index(h)
        out.append([h, units[i], vals[i]])
        d[h] = vals[i]
name = vals[hdr.index('Kernel Name')]
with open(f'profiles/{tag}_k_preprocess_tma_ncu_full.csv', 'w') as f:
    f.write(f"# {tag} ncu --set full --clock-control none --import-source on, {name}, one launch over B=20 sparse 4K frames "
            "(tools/ncu_target.py B=20 STEPS=3 -s 2 -c 1)\n")
    csv.writer(f).writerows(out)
rd, wr = float(d['dram__bytes_read.sum']) * 1e6, float(d['dram__bytes_write.sum']) * 1e6
json.dump({"kernel": name, "source": f"profiles/{tag}_k_preprocess_tma_ncu_full.csv (ncu --set full, one launch over 20 sparse 4K frames)",
           "frames_in_capture": 20, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_frame": (rd + wr) / 20},
          open(f'profiles/{tag}_k1_traffic.json', 'w'), indent=1)
for k in ('gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
          'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__block_size',
          'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
          'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio'):
    print(k, d.get(k))
rows = [r for r in csv.reader(open(launches_csv)) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    nm = r[ix['Kernel Name']].split('(')[0].replace('void ', '')
    a = agg.setdefault(nm, [0, 0.0, 0.0])
    v = float(r[ix['Metric Value']].replace(',', ''))
    if r[ix['Metric Name']] == 'gpu__time_duration.sum':
        u = r[ix['Metric Unit']]
        v = v / 1000 if u in ('ns', 'nsecond') else v * 1000 if u in ('ms', 'msecond') else v
        a[0] += 1
        a[1] += v
    else:
        a[2] += v
tot, toti = sum(v[1] for v in agg.values()), sum(v[2] for v in agg.values())
with open(f'profiles/{tag}_launch_shares.csv', 'w') as f:
    f.write(f"# {tag} launch list summary: ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:^k_ -c 500 "
            "python bench.py --steps 2 --warmup 1 --cpu-frames 0 --e2e-steps 1\n")
    f.write("# (60 sparse 4K frames per step in 3 sub-batches of 20; per-launch times are cold-cache and serialised: compare SHARES, not "
            "absolutes; minst = million warp instructions per launch)\n")
    f.write("kernel,launches,total_us,avg_us,share_pct,avg_minst,inst_share_pct\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{v[0]},{v[1]:.1f},{v[1] / v[0]:.1f},{100 * v[1] / tot:.1f},{v[2] / v[0] / 1e6:.2f},{100 * v[2] / toti:.1f}\n")
print(open(f'profiles/{tag}_launch_shares.csv').read())
