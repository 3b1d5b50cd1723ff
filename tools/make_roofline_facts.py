"""Derives profiles/roofline_facts.json (read by bench.py) from the committed `ncu --set full --page raw --csv`
export of the step kernels (profiles/r2_final_ncu_raw.csv; captured with
`ncu --set full --clock-control none --import-source on -k regex:"attn8|gemm_tc_kernel|attn_l4s|ln_mod|embed_kernel|final_kernel"
 -s 20 -c 14 python tools/run_step.py` at BASELINE configs[1], B = 64, T = 1000, L = 4)."""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "profiles", "r2_final_ncu_raw.csv")


def main():
    rows = list(csv.reader(open(SRC)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        v, u = float(r[idx[name]]), units[idx[name]]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

    ker = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        key = "attn8_prep" if "attn8_prep" in name else "attn8_main" if "attn8_kernel" in name else None
        if key and key not in ker:
            ker[key] = r
    p, m = ker["attn8_prep"], ker["attn8_main"]
    traffic = sum(val(r, n) for r in (p, m) for n in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    facts = {
        "mha_t:64x1000x4": {
            "traffic": traffic,
            "traffic_breakdown": {
                "attn8_prep_kernel": {"read": val(p, "dram__bytes_read.sum"), "write": val(p, "dram__bytes_write.sum"),
                                      "ms": float(p[idx["gpu__time_duration.sum"]])},
                "attn8_kernel": {"read": val(m, "dram__bytes_read.sum"), "write": val(m, "dram__bytes_write.sum"),
                                 "ms": float(m[idx["gpu__time_duration.sum"]])},
                "algorithmic_bytes": 64 * 1000 * 4 * (1152 * 2 + 384 * 2),
            },
            "limiter": {
                "resource": "MUFU ex2 (XU pipe) together with instruction issue: 1 exponential per score, 96 MMA FLOP per exponential",
                "xu_pipe_pct": float(m[idx["sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]]),
                "issue_active_pct": float(m[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                "tensor_pipe_pct": float(m[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]),
                "unit": "% of peak sustained (ncu, attn8_kernel)",
            },
            "source": "profiles/r2_final_ncu_raw.csv (ncu --set full, attn8_prep_kernel + attn8_kernel<4,1>) via tools/make_roofline_facts.py",
        }
    }
    out = os.path.join(ROOT, "profiles", "roofline_facts.json")
    json.dump(facts, open(out, "w"), indent=1)
    print(json.dumps(facts, indent=1))


if __name__ == "__main__":
    main()
