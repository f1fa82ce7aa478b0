"""Turn the captures tools/make_profiles.sh left in gpurun_out/ into the summaries under profiles/."""
import collections, csv, json, os, subprocess, sys
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ['gpu__time_duration.sum', 'launch__block_size', 'launch__grid_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']


def summarise(rep, out, header):
    path = os.path.join(G, rep)
    if not os.path.exists(path):
        print("missing", path); return
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(P, out), "w") as f:
        f.write(header + "\n")
        for r in rows[2:]:
            d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
            f.write("=====\nKernel Name = %s\n" % d['Kernel Name'])
            for k in KEYS:
                if k in d and d[k] != '':
                    f.write("%s [%s] = %s\n" % (k, u[k], d[k]))
    print("wrote", out)


def launches(csv_in, md_out, bench_log):
    rows = [r for r in csv.reader(open(os.path.join(G, csv_in))) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
    H, data = rows[h], rows[h + 1:]
    ik, iv, iu = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in data:
        v = float(r[iv].replace(',', ''))
        v = v / 1000 if r[iu] == 'ns' else v * 1000 if r[iu] == 'ms' else v
        agg.setdefault(r[ik].split('(')[0].replace('void ', '').replace('<unnamed>::', ''), []).append(v)
    bench = json.loads(open(os.path.join(G, bench_log)).read().strip().splitlines()[-1])
    frame = ['band_list_kernel', 'raymarch_persistent<0, 768, 0>', 'retrace_kernel<0>']
    frame += [k for k in agg if k.startswith(('bloom_h', 'bloom_v', 'composite_kernel', 'flare_add'))]
    med = lambda v: sorted(v)[len(v) // 2]      # (median: the banded synchronous frames of the e2e block launch the same kernels on row bands)
    tot = sum(med(agg[k]) for k in frame if k in agg)
    with open(os.path.join(P, md_out), "w") as f:
        f.write(f"# {TAG} launch list summary (profiles/{csv_in})\n\n")
        f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --no-orbit --no-tiled`\n"
                f"(after the same command exited 0 without ncu: ms_per_step {bench['ms_per_step']:.4f}, stage_ms {bench['stage_ms']}).\n"
                "Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's stage_ms, not absolutes.\n\n")
        f.write("| kernel | launches | median us | share of one frame |\n|---|---|---|---|\n")
        for k, v in agg.items():
            m = med(v)
            share = f"{100 * m / tot:.1f} %" if k in frame else "(outside the timed frame: setup / orbit-video block / peak probe)"
            f.write(f"| {k} | {len(v)} | {m:.1f} | {share} |\n")
        st = bench['stage_ms']; s = sum(st.values())
        rm = sum(med(agg[k]) for k in frame[:3] if k in agg)
        f.write(f"\nOne frame under ncu = {tot:.1f} us; ray march (band list + persistent + retrace) = {100 * rm / tot:.1f} % of it; "
                f"bench.py stage_ms (CUDA events, no profiler): ray march {100 * st['ray_march'] / s:.1f} %, bloom H {100 * st['bloom_h'] / s:.1f} %, "
                f"bloom V + composite {100 * st['bloom_v_composite'] / s:.1f} %.\n")
    print("wrote", md_out)


if __name__ == "__main__":
    import shutil
    shutil.copy(os.path.join(G, f"{TAG}_launches.csv"), os.path.join(P, f"{TAG}_launches.csv"))
    launches(f"{TAG}_launches.csv", f"{TAG}_launches_summary.md", f"{TAG}_bench_plain.log")
    summarise(f"{TAG}_raymarch.ncu-rep", f"{TAG}_raymarch_ncu.txt", "ncu --set full --clock-control none: fhd default scene (tools/prof_fhd.py fhd 0 3), ray-march kernels")
    summarise(f"{TAG}_post.ncu-rep", f"{TAG}_post_ncu.txt", "ncu --set full --clock-control none: fhd default scene, bloom H / bloom V / composite")
    summarise(f"{TAG}_raymarch_4k_aa.ncu-rep", f"{TAG}_raymarch_4k_aa_ncu.txt", "ncu --set full --clock-control none: 4K, anti_alias lod_radius, tilt 20, flare (BASELINE configs[2]) ray march with differentials")
    summarise(f"{TAG}_png.ncu-rep", f"{TAG}_png_ncu.txt", "ncu --set full --clock-control none: device PNG encoder (tools/prof_png.py), fhd default scene frame")
    summarise(f"{TAG}_texture.ncu-rep", f"{TAG}_texture_ncu.txt", "ncu --set full --clock-control none: disk-texture pipeline kernels of a video frame (tools/video_breakdown.py)")
