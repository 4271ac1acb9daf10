"""Files the output of tools/gpu_round.sh (gpurun_out/, scratch) under profiles/ (tracked) and derives the summaries
bench.py, DESIGN.md and the judge cite.

usage: python tools/refresh_profiles.py [tag, default r2]
  profiles/bench_<tag>_n1.json, bench_<tag>_reference_arm.json   the two bench lines
  profiles/<tag>_ops_event_timings.txt                           per-op CUDA-event timings of one denoiser step
  profiles/<tag>_launchlist_ncu.csv / _summary.csv               ncu launch list of the LAST denoiser step of the profiled
                                                                 program (durations, DRAM bytes) and its per-kernel shares
  profiles/<tag>_ncu_full_<capture>_raw.csv, <tag>_ncu_full_summary.json   the --set full captures (raw page + key metrics)
  profiles/<tag>_sass_opcodes.txt                                tcgen05 / TMA / TMEM opcode histogram of the shipped library
"""
import collections
import csv
import glob
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
MULT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6,
        "nsecond": 1.0}


def last_json(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def short(name):
    return name.split("(")[0].replace("void ", "").replace("cm::", "").replace("(anonymous namespace)::", "")


def bench_lines():
    for src, dst in (("bench.json", f"bench_{tag}_n1.json"), ("bench_ref.json", f"bench_{tag}_reference_arm.json")):
        p = os.path.join(G, src)
        if os.path.exists(p):
            json.dump(last_json(p), open(os.path.join(P, dst), "w"))
    if os.path.exists(os.path.join(G, "ops.txt")):
        shutil.copy(os.path.join(G, "ops.txt"), os.path.join(P, f"{tag}_ops_event_timings.txt"))


def launch_list():
    path = os.path.join(G, "launches.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    by = collections.OrderedDict()
    for r in rows:                                   # one row per (launch id, metric)
        d = by.setdefault(int(r[0]), {"name": short(r[4]), "grid": r[8], "block": r[7]})
        d[r[-3]] = float(r[-1].replace(",", "")) * MULT.get(r[-2], 1.0)
    launches = list(by.values())
    finals = [i for i, d in enumerate(launches) if d["name"].startswith("final_conv_kernel")]
    if len(finals) >= 2:                             # the last denoiser step: after the previous final conv, up to the last
        launches = launches[finals[-2] + 1:finals[-1] + 1]
    with open(os.path.join(P, f"{tag}_launchlist_ncu.csv"), "w") as f:
        f.write("# ncu launch list of the last denoiser step of tools/profile_ops.py (ATC B=64; --clock-control none "
                "--cache-control none): duration ns, DRAM bytes read / written\n# command: tools/gpu_round.sh stage 2\n"
                "index,kernel,grid,block,duration_ns,dram_read_bytes,dram_write_bytes\n")
        for i, d in enumerate(launches):
            f.write(f"{i},\"{d['name']}\",\"{d['grid']}\",\"{d['block']}\",{d.get('gpu__time_duration.sum', 0):.0f},"
                    f"{d.get('dram__bytes_read.sum', 0):.0f},{d.get('dram__bytes_write.sum', 0):.0f}\n")
    agg = collections.OrderedDict()
    for d in launches:
        a = agg.setdefault(d["name"], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values()) or 1.0
    with open(os.path.join(P, f"{tag}_launchlist_summary.csv"), "w") as f:
        f.write("# per-kernel shares of one denoiser step, ATC B=64 (ncu gpu__time_duration.sum: cold-cache, serialised launches "
                "-- the SHARES are what compares with the CUDA-event numbers of bench.py)\nkernel,launches,total_ns,share_of_step,dram_bytes\n")
        for k, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{n},{t:.0f},{t / tot:.3f},{b:.0f}\n")
        f.write(f"# total,{len(launches)},{tot:.0f},1.000\n")


WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def full_captures():
    summ = {}
    for raw in sorted(glob.glob(os.path.join(G, "full_*_raw.csv"))):
        name = os.path.basename(raw)[len("full_"):-len("_raw.csv")]
        rr = list(csv.reader(open(raw)))
        if len(rr) < 3:
            continue
        shutil.copy(raw, os.path.join(P, f"{tag}_ncu_full_{name}_raw.csv"))
        hdr, units = rr[0], rr[1]
        for r in rr[2:]:
            d = {}
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    d[w] = (r[i] + " " + units[i]).strip()
            summ.setdefault(name, []).append(d)
    if summ:
        json.dump(summ, open(os.path.join(P, f"{tag}_ncu_full_summary.json"), "w"), indent=1)


def sass_histogram():
    lib = os.path.join(ROOT, "crowdmod-ddpm-4d_b200", "libcrowdmod_b200.so")
    if not os.path.exists(lib):
        return
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    per_kernel = collections.OrderedDict()
    cur = None
    pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCCP|UTCBAR|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UBLKCP|SYNCS|HMMA|QGMMA|UCGABAR_ARV|UCGABAR_WAIT)(\.[\w.]+)?")
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = pat.search(line)
        if m and cur:
            per_kernel.setdefault(cur, collections.Counter())[m.group(1) + (m.group(2) or "")] += 1
    total = collections.Counter()
    for c in per_kernel.values():
        total.update(c)
    demangle = {}
    try:
        names = list(per_kernel)
        dm = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    except Exception:  # noqa: BLE001
        pass
    with open(os.path.join(P, f"{tag}_sass_opcodes.txt"), "w") as f:
        f.write("# cuobjdump -sass crowdmod-ddpm-4d_b200/libcrowdmod_b200.so (sm_100a): tensor-core / TMEM / TMA / mbarrier opcodes\n")
        f.write("# UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA load), UBLKCP = cp.async.bulk, "
                "SYNCS = mbarrier, HMMA = mma.sync\n\n## whole library\n")
        for k, v in sorted(total.items(), key=lambda kv: -kv[1]):
            f.write(f"{v:6d}  {k}\n")
        f.write("\n## per kernel (kernels with tcgen05 / TMA instructions)\n")
        for k, c in per_kernel.items():
            if not any(o.startswith(("UTC", "UTMA", "LDTM", "UBLKCP")) for o in c):
                continue
            f.write(f"\n{short(demangle.get(k, k))[:110]}\n")
            for o, v in sorted(c.items(), key=lambda kv: -kv[1]):
                f.write(f"    {v:5d}  {o}\n")


if __name__ == "__main__":
    os.makedirs(P, exist_ok=True)
    bench_lines()
    launch_list()
    full_captures()
    sass_histogram()
    print("profiles/ refreshed for", tag)
