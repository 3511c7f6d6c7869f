"""Where the chain kernels wait, from an .ncu-rep captured with `ncu --set full --import-source on`:
    python tools/ncu_stalls.py gpurun_out/prof_r2_step.ncu-rep profiles/r02_chain_stalls_l2.txt
Part 1 (raw page): L2 -> SM and SM -> L2 bytes per launch (l1tex__m_xbar2l1tex_read_bytes / l1tex__m_l1tex2xbar_write_bytes), next to
the duration, DRAM bytes and tensor-pipe activity: the weight chunks every CTA re-streams for every 128-point tile travel this path.
Part 2 (source page, SASS view): the instructions that collect the most warp-stall samples per kernel; mbarrier try-wait loops are
labelled with the barrier they poll (offset inside the barrier block of chain_common.cuh's `Bars`)."""
import csv, io, re, subprocess, sys

KERNELS = ("sdf_fused_kernel", "color_fused_kernel", "sdf_chain_query", "tc_wgrad_kernel")
BAR_NAMES = [(0x00, "w_full"), (0x40, "w_empty"), (0x80, "aux_full"), (0xC0, "aux_empty"), (0x100, "a_ready"), (0x140, "acc_full"),
             (0x150, "stg_full"), (0x160, "stg_empty"), (0x170, "a_free"), (0x178, "a_init"), (0x180, "tile_done"),
             (0x188, "h_stored"), (0x190, "epi_done")]


def ncu_page(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def short(name):
    return name.replace("(int)", "").split("(")[0].replace("void ", "").replace("cope::", "").replace("<unnamed>::", "").replace("unnamed>::", "")


def bar_name(off, base):
    rel = off - base
    best = None
    for o, n in BAR_NAMES:
        if rel >= o:
            best = (o, n)
    return f"{best[1]}+{rel - best[0]:#x}" if best and 0 <= rel < 0x200 else f"smem+{off:#x}"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    lines = []
    rows = list(csv.reader(io.StringIO(ncu_page(rep, "raw"))))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, name, scale=1.0):
        try:
            u = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "us": 1, "ms": 1e3, "ns": 1e-3}.get(units[ix[name]], 1)
            return float(r[ix[name]].replace(",", "")) * u * scale
        except Exception:
            return float("nan")

    lines.append("# part 1: bytes through the L2 <-> SM path per launch (1024-ray training step, 131 072 points)")
    lines.append(f"{'kernel':28s} {'us':>7s} {'L2->SM GB':>10s} {'SM->L2 GB':>10s} {'L2<->SM TB/s':>13s} {'DRAM GB':>8s} {'lts thr %':>9s} {'tensor %':>9s}")
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        if not any(k in name for k in KERNELS):
            continue
        t = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "l1tex__m_xbar2l1tex_read_bytes.sum"), val(r, "l1tex__m_l1tex2xbar_write_bytes.sum")
        dram = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        lines.append(f"{short(name):28s} {t:7.1f} {rd / 1e9:10.3f} {wr / 1e9:10.3f} {(rd + wr) / t / 1e6:13.2f} {dram / 1e9:8.3f} "
                     f"{val(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):9.1f} "
                     f"{val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):9.1f}")

    src = ncu_page(rep, "source").splitlines()
    starts = [i for i, l in enumerate(src) if l.startswith('"Kernel Name"')] + [len(src)]
    seen = set()
    lines.append("")
    lines.append("# part 2: top warp-stall sample sites per kernel (SASS view; samples of all 20 warps of a CTA, the 16 epilogue warps dominate)")
    for a, b in zip(starts[:-1], starts[1:]):
        name = next(csv.reader([src[a]]))[1]
        if not any(k in name for k in KERNELS) or name in seen:
            continue
        body = list(csv.reader(src[a + 1:b]))
        h = body[0]
        if "# Samples" not in h:
            continue
        si, ai = h.index("# Samples"), h.index("Source")
        ins = [r for r in body[1:] if len(r) > si and r[si].isdigit()]
        if not ins or not re.match(r"\s*(@!?U?P\d+\s+)?[A-Z0-9_.]+(\s|$)", ins[0][ai]):
            continue                      # the CUDA-C view of the same kernel follows the SASS view
        seen.add(name)
        tot = sum(int(r[si]) for r in ins) or 1
        # barrier block base: the lowest try-wait offset rounded down to 0x100 is w_full of the carve-up
        offs = [int(m.group(1), 16) for r in ins for m in [re.search(r"TRYWAIT.*\+0x([0-9a-f]+)\]", r[ai])] if m]
        base = (min(offs) // 0x100) * 0x100 if offs else 0
        lines.append(f"## {short(name)}: {tot} samples")
        agg = {}
        for k, r in enumerate(ins):
            s = int(r[si])
            text = r[ai].strip()
            if "BRA" in text and k > 0 and "TRYWAIT" in ins[k - 1][ai]:      # the loop branch belongs to its try-wait
                text = ins[k - 1][ai].strip()
            m = re.search(r"TRYWAIT.*\+0x([0-9a-f]+)\]", text)
            if m and "_fused_" in name:           # the two fused families share chain_common.cuh's barrier block
                key = f"mbarrier wait {bar_name(int(m.group(1), 16), base)}"
            elif m:
                key = f"mbarrier wait [barrier block + {int(m.group(1), 16) - base:#x}]"
            else:
                key = re.sub(r"\s+", " ", text)[:60]
            agg[key] = agg.get(key, 0) + s
        for key, s in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
            lines.append(f"   {100.0 * s / tot:5.1f} %  {key}")
    open(out, "w").write("\n".join(lines) + "\n")
    print(f"{out}: {len(seen)} kernels")


if __name__ == "__main__":
    main()
