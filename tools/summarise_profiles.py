"""Turn the ncu artefacts of a round into the tracked summaries under profiles/.

    ncu -i gpurun_out/prof_<tag>_sparse.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/summarise_profiles.py <tag> /tmp/raw.csv gpurun_out/<tag>_launches.csv

writes profiles/<tag>_preprocess_kernels_ncu_full.csv (one column per captured kernel), profiles/<tag>_traffic.json
(DRAM bytes per frame per kernel id, read by bench.py's roofline.traffic) and profiles/<tag>_launch_shares.csv.
"""
import collections, csv, json, sys

tag, raw_csv, launches_csv = sys.argv[1:4]
FRAMES = 20
rows = list(csv.reader(open(raw_csv)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__block_size', 'launch__grid_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'launch__waves_per_multiprocessor']
want += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
kcol = hdr.index('Kernel Name')
kernels = rows[2:]
names = [r[kcol].split('(')[0].replace('void ', '') for r in kernels]
with open(f'profiles/{tag}_preprocess_kernels_ncu_full.csv', 'w') as f:
    f.write(f"# {tag}: ncu --set full --clock-control none --import-source on -k regex:'k_preprocess_tma|k_sparse' -s 3 -c 3, one launch each over "
            f"B={FRAMES} sparse 4K frames (tools/ncu_target.py B=20 STEPS=2); k_preprocess_tma<0,40,3840,1,1,3> = the bounds pass (MODE 1)\n")
    wr = csv.writer(f)
    wr.writerow(['metric', 'unit'] + names)
    for h in want:
        if h in hdr:
            i = hdr.index(h)
            wr.writerow([h, units[i]] + [r[i] for r in kernels])
traffic = {}
kid = {'k_preprocess_tma': 'k_preprocess_fused', 'k_sparse_flags': 'k_sparse_flags', 'k_sparse_exact': 'k_sparse_exact'}
for r, n in zip(kernels, names):
    rd = float(r[hdr.index('dram__bytes_read.sum')]) * 1e6
    wrb = float(r[hdr.index('dram__bytes_write.sum')]) * 1e6
    key = next(v for k, v in kid.items() if n.startswith(k))
    traffic[key] = {"kernel": n, "frames_in_capture": FRAMES, "dram_bytes_read": rd, "dram_bytes_write": wrb,
                    "dram_bytes_per_frame": (rd + wrb) / FRAMES, "time_us": float(r[hdr.index('gpu__time_duration.sum')])}
json.dump(traffic, open(f'profiles/{tag}_traffic.json', 'w'), indent=1)

# launch list -> shares of the serial step (second step of the capture)
lrows = list(csv.reader(open(launches_csv)))
hi = [i for i, r in enumerate(lrows) if 'Kernel Name' in r][0]
h2 = lrows[hi]
k, m, v, idc = h2.index('Kernel Name'), h2.index('Metric Name'), h2.index('Metric Value'), h2.index('ID')
per = collections.OrderedDict()
for r in lrows[hi + 1:]:
    if len(r) > v:
        per.setdefault((int(r[idc]), r[k].split('(')[0].replace('void ', '')[:48]), {})[r[m]] = float(r[v].replace(',', ''))
ids = list(per)
half = [i for i in ids if i[0] >= ids[len(ids) // 2][0]]
agg = collections.OrderedDict()
for i in half:
    d = per[i]
    a = agg.setdefault(i[1], [0.0, 0.0, 0, 0.0])
    a[0] += d['gpu__time_duration.sum'] / 1e3
    a[1] += d['smsp__inst_executed.sum'] / 1e6
    a[2] += 1
    a[3] = d['smsp__issue_active.avg.pct_of_peak_sustained_active']
tt, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
with open(f'profiles/{tag}_launch_shares.csv', 'w') as f:
    f.write(f"# {tag}: ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active --clock-control none, tools/ncu_target.py B=20 STEPS=2, "
            "second step (serialised, cold cache: shares, not absolutes)\n")
    wr = csv.writer(f)
    wr.writerow(['kernel', 'launches', 'time_us', 'time_share', 'warp_instructions_M', 'instruction_share', 'issue_active_pct_last_launch'])
    for n, a in agg.items():
        wr.writerow([n, a[2], round(a[0], 1), round(a[0] / tt, 4), round(a[1], 2), round(a[1] / ti, 4), round(a[3], 1)])
    wr.writerow(['total', sum(a[2] for a in agg.values()), round(tt, 1), 1.0, round(ti, 2), 1.0, ''])
print(open(f'profiles/{tag}_launch_shares.csv').read())
print(json.dumps(traffic, indent=1))
