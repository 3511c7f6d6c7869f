"""Summarise an .ncu-rep (ncu --set full) into the per-launch CSV committed under profiles/:
    python tools/ncu_summary.py gpurun_out/prof_r2_step.ncu-rep profiles/r02_step_kernels_ncu_full.csv [mlp]
Columns: kernel, grid, duration_us, dram_read_bytes, dram_write_bytes, dram_gbps, tensor_pipe_pct, issue_active_pct, registers,
mlp_group (1 for the kernels bench.py's roofline.traffic sums: chain / wgrad / colour / query kernels)."""
import csv, io, subprocess, sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "sm__warps_active.avg.pct_of_peak_sustained_active"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "usecond": 1, "nsecond": 1e-3,
        "msecond": 1e3, "second": 1e6}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        if name not in ix or r[ix[name]] in ("", "n/a"):
            return None
        v = float(r[ix[name]].replace(",", ""))
        return v * UNIT.get(units[ix[name]], 1)

    mlp_names = ("sdf_fused", "tc_wgrad", "color_fused", "sdf_chain")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "grid", "duration_us", "dram_read_bytes", "dram_write_bytes", "dram_gbps", "tensor_pipe_pct",
                    "issue_active_pct", "warps_active_pct", "registers", "mlp_group"])
        tot = [0.0, 0.0, 0.0]
        for r in rows[2:]:
            name = r[ix["Kernel Name"]]
            short = name.split("(")[0].replace("void ", "").replace("cope::", "").replace("<unnamed>::", "")
            t, rd, wr = val(r, METRICS[0]), val(r, METRICS[1]), val(r, METRICS[2])
            tp = val(r, METRICS[3])
            if tp is None:
                tp = val(r, METRICS[4])
            mlp = int(any(k in name for k in mlp_names))
            if mlp:
                tot[0] += t; tot[1] += rd; tot[2] += wr
            w.writerow([short, r[ix["Grid Size"]].replace(",", " "), f"{t:.1f}", f"{rd:.0f}", f"{wr:.0f}", f"{(rd + wr) / t / 1e3:.0f}",
                        "" if tp is None else f"{tp:.1f}", f"{val(r, METRICS[5]):.1f}", f"{val(r, METRICS[7]):.1f}",
                        f"{val(r, METRICS[6]):.0f}", mlp])
    print(f"{out}: MLP group {tot[0]:.0f} us, {(tot[1] + tot[2]) / 1e9:.2f} GB of DRAM traffic ({tot[1] / 1e9:.2f} read + {tot[2] / 1e9:.2f} written)")


if __name__ == "__main__":
    main()
