"""Copies the round's evidence from gpurun_out/ (scratch) into profiles/ (tracked) and derives the summaries
bench.py and DESIGN.md cite: launch-list shares, the full-set ncu capture of conv_plane_kernel, DRAM traffic."""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

def last_json(path):
    return json.loads(open(path).read().strip().splitlines()[-1])

json.dump(last_json(os.path.join(G, "bench.json")), open(os.path.join(P, f"bench_{tag}_n1.json"), "w"))
json.dump(last_json(os.path.join(G, "bench_ref.json")), open(os.path.join(P, f"bench_{tag}_reference_arm.json"), "w"))
shutil.copy(os.path.join(G, "ops.txt"), os.path.join(P, f"{tag}_ops_event_timings.txt"))
shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, f"{tag}_launchlist_ncu.csv"))
rows = [r for r in csv.reader(open(os.path.join(G, "launches.csv"))) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r[4].split("(")[0].replace("void ", ""), [0, 0.0])
    a[0] += 1
    a[1] += float(r[-1].replace(",", ""))
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, f"{tag}_launchlist_summary.csv"), "w") as f:
    f.write("# ncu launch list summary, one denoiser step, ATC B=64 (gpu__time_duration.sum, ns; --clock-control none --cache-control none)\n")
    f.write("# command: see tools/gpu_round.sh\nkernel,launches,total_ns,share_of_step\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{n},{t:.0f},{t / tot:.3f}\n")
    f.write(f"# total,{len(rows)},{tot:.0f},1.000\n")
rep = os.path.join(G, "conv_plane_full.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(P, f"{tag}_conv_plane_ncu_full_raw.csv"), "w").write(raw)
rr = list(csv.reader(raw.splitlines()))
hdr, units, data = rr[0], rr[1], rr[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg"]
launches, dram = [], []
mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in data:
    d = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            d[w] = (r[i] + " " + units[i]).strip()
    launches.append(d)
    b = 0.0
    for w in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(w)
        b += float(r[i].replace(",", "")) * mult.get(units[i], 1.0)
    dram.append(b)
json.dump({"command": "ncu --set full --clock-control none --import-source on -k regex:conv_plane -s 33 -c 3 python tools/profile_ops.py 64 (tools/gpu_round.sh)",
           "note": "ATC B=64, one denoiser step; three consecutive conv_plane_kernel launches of the decoder's full-resolution blocks",
           "launches": launches}, open(os.path.join(P, f"{tag}_conv_plane_ncu_full_summary.json"), "w"), indent=1)
json.dump({"kernel": "conv_plane_kernel", "dram_bytes_per_launch": sum(dram) / len(dram), "per_launch": dram,
           "source": f"profiles/{tag}_conv_plane_ncu_full_summary.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"},
          open(os.path.join(P, "conv_umma_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, f"{tag}_launchlist_summary.csv")).read())
print(json.dumps(launches, indent=0)[:1800])
