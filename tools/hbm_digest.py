"""HBM evidence for the bandwidth-bound kernels: per captured launch of an .ncu-rep (ncu --set full) print the duration,
the DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and the achieved GB/s against the measured copy peak
(MEASURED_PEAKS.json: 6531.9 GB/s) and the ~8 TB/s spec.   usage: hbm_digest.py file.ncu-rep"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def num(s):
    return float(s.replace(",", ""))


def main(path):
    peak = 6531.9
    pj = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pj):
        peak = json.load(open(pj))["hbm_gbs"]
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    unit = dict(zip(h, units))
    print(f"{'kernel':42s} {'grid':>6s} {'us':>8s} {'DRAM MB':>9s} {'GB/s':>8s} {'% of measured':>14s} {'% of 8 TB/s':>12s}  L2->SM MB")
    for r in rows[2:]:
        d = dict(zip(h, r))
        def val(k, scale_units):
            v = num(d[k]); u = unit[k]
            return v * scale_units.get(u, 1.0)
        t_us = val("gpu__time_duration.sum", {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3})
        byt = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd, wr = val("dram__bytes_read.sum", byt), val("dram__bytes_write.sum", byt)
        l2 = val("l1tex__m_xbar2l1tex_read_bytes.sum", byt) if "l1tex__m_xbar2l1tex_read_bytes.sum" in d else 0.0
        gbs = (rd + wr) / t_us / 1e3
        name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("unnamed>::", "").replace("destr::", "")[:42]
        print(f"{name:42s} {d['launch__grid_size']:>6s} {t_us:8.2f} {(rd + wr) / 1e6:9.2f} {gbs:8.0f} {100 * gbs / peak:13.1f}% {100 * gbs / 8000:11.1f}%  {l2 / 1e6:8.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
