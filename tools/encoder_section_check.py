"""Reads an ncu launch list (csv, one training step, launch order) and reports the library GEMM kernels (nvjet_* /
cutlass* / sgemm) that fall inside the ENCODER parts of the step:
  forward  = launches up to the encoder's last kernel (the last add_ln2_fwd_kernel = norm2 + shared norm of layer L-1);
             what follows -- fine_pos, the hoisted decoder key/value/position projections -- is the decoder's input side
  backward = launches from the first add_ln2_bwd_kernel (backward of layer L-1's tail) to the optimizer.
Kernels on forked streams interleave in launch order, so the side-stream weight-gradient kernels appear in between."""
import csv, re, sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
names = [re.sub(r"\(.*", "", r["Kernel Name"]) for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
lib = lambda n: n.startswith("nvjet") or "cutlass" in n or "gemv" in n or "sgemm" in n
idx = lambda pat: [i for i, n in enumerate(names) if pat in n]
f_end = max(idx("add_ln2_fwd_kernel"))
b_start = min(idx("add_ln2_bwd_kernel"))
print(f"{len(names)} launches in the step; library GEMM kernels in the whole step: {sum(lib(n) for n in names)}")
for label, lo, hi in (("encoder forward ", 0, f_end + 1), ("encoder backward", b_start, len(names))):
    inside = [(i, names[i]) for i in range(lo, hi) if lib(names[i])]
    ours = sum(1 for i in range(lo, hi) if "destr::" in names[i])
    print(f"  {label}: launches [{lo}, {hi}): {hi - lo} kernels, {ours} of them destr::, library GEMM kernels inside: {len(inside)}")
    for i, n in inside:
        print(f"      #{i}  {n[:90]}")
gemm = [n for n in names[:f_end + 1] + names[b_start:] if "gemm_" in n]
print(f"  tcgen05 GEMM-family launches inside the two encoder parts: {len(gemm)}")
