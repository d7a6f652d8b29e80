"""Summarise an .ncu-rep (from `ncu --set full`) into JSON: per kernel launch — duration, DRAM bytes, DRAM / tensor-pipe /
XU (MUFU) / shared-memory utilisation, registers.  Usage: python tools/ncu_summary.py REPORT.ncu-rep [labels,comma,separated] > out.json"""
import csv
import io
import json
import subprocess
import sys

WANT = {   # exact column names of `ncu --page raw --csv`
    "duration_us": "gpu__time_duration.sum",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "l1tex_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smem_tc_wavefronts_pct": "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smem_lsu_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm_clock_ghz": "sm__cycles_elapsed.avg.per_second",
    "regs": "launch__registers_per_thread",
}
UNIT_SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
              "Ghz": 1.0, "Mhz": 1e-3}


def main():
    rep = sys.argv[1]
    labels = sys.argv[2].split(",") if len(sys.argv) > 2 else []
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]

    cols = {k: (hdr.index(v) if v in hdr else None) for k, v in WANT.items()}
    out = []
    for n, r in enumerate(data):
        d = {"kernel": labels[n] if n < len(labels) else "", "name": r[hdr.index("Kernel Name")][:90],
             "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
        for k, i in cols.items():
            if i is None or r[i] in ("", "n/a"):
                d[k] = None
                continue
            v = float(r[i].replace(",", ""))
            d[k] = v * UNIT_SCALE.get(units[i].split("/")[0], 1.0)
        if d.get("dram_read_bytes") is not None and d.get("dram_write_bytes") is not None:
            d["dram_bytes"] = d["dram_read_bytes"] + d["dram_write_bytes"]
        out.append(d)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
