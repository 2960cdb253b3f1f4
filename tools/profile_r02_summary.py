"""Turns the captures of tools/profile_r02.sh (gpurun_out/r02prof/) into the summaries committed under profiles/ (round 2).
Reads the .ncu-rep files with `ncu -i ... --page raw --csv` (no GPU needed)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "r02prof")
OUT = os.path.join(ROOT, "profiles")

WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/shared throughput % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
        ("smsp__issue_active.avg.pct", "issue slots active %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "registers / thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"), ("sm__cycles_elapsed.max", "elapsed cycles")]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def summarise(name, title):
    rep = os.path.join(SRC, name + ".ncu-rep")
    if not os.path.exists(rep):
        return None, f"# {title}\n(capture missing)\n"
    hdr, units, rows = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    out = [f"# {title}", f"# ncu --set full --clock-control none --import-source on, round-2 final build; source: gpurun_out/r02prof/{name}.ncu-rep"]
    recs = []
    for r in rows:
        out.append(f"kernel: {r[idx['Kernel Name']]}")
        rec = {"kernel": r[idx["Kernel Name"]]}
        for key, label in WANT:
            if key in idx:
                out.append(f"  {label:42s} {r[idx[key]]} {units[idx[key]]}")
                rec[key] = (r[idx[key]], units[idx[key]])
        st = []
        for h in hdr:
            if "smsp__average_warps_issue_stalled" in h and "_per_issue_active" in h and "not_issued" not in h:
                try:
                    st.append((float(r[idx[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        st.sort(reverse=True)
        out.append("  stalls per issue (top): " + ", ".join(f"{n} {v:.2f}" for v, n in st[:5]))
        recs.append(rec)
        out.append("")
    return recs, "\n".join(out) + "\n"


def to_bytes(v, unit):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    os.makedirs(OUT, exist_ok=True)
    texts = []
    traffic = None
    for name, title in [("gemm_pair_6016x1280x10240", "tcgen05 GEMM, CTA-pair mode (cta_group::2, block_n = 256 forced), 6016 x 1280 x 10240, bf16 out"),
                        ("gemm_auto_6016x1280x10240", "tcgen05 GEMM, the library's own tile choice for 6016 x 1280 x 10240 (224-wide, one CTA per SM), bf16 out"),
                        ("gemm_shortk_24064x2560x320", "tcgen05 GEMM, short contraction, 24064 x 2560 x 320 (FF1 of level 0), bf16 out"),
                        ("gemm_wgrad_1280x1280x6016", "tcgen05 GEMM, weight gradient (stream-K, fp32 vector-RED flush), 1280 x 1280 x 6016")]:
        recs, txt = summarise(name, title)
        texts.append(txt)
        if recs and name.startswith("gemm_pair"):
            r = recs[0]
            M, N, K = 6016, 1280, 10240
            traffic = {"kernel": "gemm_kernel<256, 2, true> (CTA pair)", "shape": f"M={M} N={N} K={K}, bf16 out, K-major A and B",
                       "dram_bytes_read": to_bytes(*r["dram__bytes_read.sum"]), "dram_bytes_write": to_bytes(*r["dram__bytes_write.sum"]),
                       "algorithmic_bytes": 2.0 * (M * K + N * K + M * N),
                       "source": "profiles/r02_gemm_full_summary.txt (ncu --set full --clock-control none, one launch, round-2 final build)"}
    open(os.path.join(OUT, "r02_gemm_full_summary.txt"), "w").write("\n".join(texts))
    if traffic:
        json.dump(traffic, open(os.path.join(OUT, "r02_roofline_traffic.json"), "w"), indent=1)
    _, txt = summarise("attn_d40_self", "fused tcgen05 attention, level-0 self-attention (32 x 8 heads x 752 x 752, d = 40): forward, dQ, dK/dV")
    open(os.path.join(OUT, "r02_attention_full_summary.txt"), "w").write(txt)
    _, t1 = summarise("rvq_tc", "RVQ quantise, tcgen05 pre-selection + exact fp32 re-ranking, 256 clips x 900 frames")
    _, t2 = summarise("rvq_rest", "RVQ: exhaustive exact-fp32 quantiser and the two embedding-sum kernels (32 clips x 900 / 512 clips x 900)")
    open(os.path.join(OUT, "r02_rvq_full_summary.txt"), "w").write(t1 + "\n" + t2)
    # launch list
    la = os.path.join(SRC, "launches_all.csv")
    ll = os.path.join(SRC, "launches_step.csv")
    if os.path.exists(la):      # keep the last step: from the last add_noise kernel (once per step) to the end
        lines = [l for l in open(la) if not l.startswith("==")]
        last = max(i for i, l in enumerate(lines) if "add_noise" in l)
        open(ll, "w").writelines([lines[0]] + lines[last:])
    if os.path.exists(ll):
        lines = [l for l in open(ll) if not l.startswith("==")]
        open(os.path.join(OUT, "r02_launches_step.csv"), "w").writelines(lines)
        s = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), ll, "12"], capture_output=True, text=True).stdout
        head = ("# ncu --metrics gpu__time_duration.sum --clock-control none -s 5500, `python bench.py --no-graph --steps 2 --warmup 3 "
                "--no-cpu-baseline --no-sampling --no-rvq --no-full-step`: the last (steady) eager train step of the bench workload, round-2 final build.\n"
                "# Per-launch times are cold-cache and serialised: compare SHARES with the in-graph figures of tools/step_profile.py.\n")
        open(os.path.join(OUT, "r02_launches_step_summary.txt"), "w").write(head + s)
    sm = os.path.join(SRC, "sass_mnemonics.txt")
    if os.path.exists(sm):
        open(os.path.join(OUT, "r02_sass_mnemonics.txt"), "w").write(
            "# cuobjdump -sass prompt_tts_b200/libpt_b200.so | grep -o <mnemonic> | sort | uniq -c   (round-2 final build)\n" + open(sm).read())
    print("written:", sorted(f for f in os.listdir(OUT) if f.startswith("r02_")))


if __name__ == "__main__":
    main()
