"""SASS evidence for profiles/: per kernel of liblfpsqp_b200.so the instruction count and the counts of the mnemonics that
identify the data path (DMMA = mma.sync f64, DFMA, LDGSTS = cp.async, SHFL, BAR, UTMALDG / UTCMMA = TMA / tcgen05 -- absent:
there is no FP64 tcgen05 kind), plus the full listing of the DMMA GEMM instantiations.
   python tools/sass_summary.py [out_prefix]      (needs cuobjdump; runs without a GPU)"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "lfpsqp.jl_b200", "liblfpsqp_b200.so")
prefix = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "profiles", "r2_sass")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
watch = ["DMMA", "DFMA", "DADD", "DMUL", "MUFU", "LDGSTS", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "ATOM", "RED", "UTMALDG",
         "UTMASTG", "UTCMMA", "HMMA", "LDL", "STL"]
kern = None; counts = collections.OrderedDict(); listing = collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)[:140]
        counts[kern] = collections.Counter(); listing[kern] = []
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if kern and m:
        op = m.group(2).split(".")[0]
        counts[kern]["_all"] += 1
        if op in watch: counts[kern][op] += 1
        listing[kern].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line.rstrip()))   # drop the encoding column
with open(prefix + "_summary.txt", "w") as f:
    f.write("cuobjdump -sass lfpsqp.jl_b200/liblfpsqp_b200.so (sm_100a): instructions per kernel and data-path mnemonics\n")
    f.write("(UTMALDG / UTCMMA = 0 everywhere: the FP64 tensor path on sm_100a is mma.sync m8n8k4 = DMMA; cp.async = LDGSTS)\n\n")
    tot = collections.Counter()
    for k, c in counts.items():
        f.write("%-140s %6d  %s\n" % (k, c["_all"], " ".join("%s=%d" % (w, c[w]) for w in watch if c[w])))
        tot.update(c)
    f.write("\nTOTAL %d  %s\n" % (tot["_all"], " ".join("%s=%d" % (w, tot[w]) for w in watch)))
with open(prefix + "_dgemm_nt_kernel.txt", "w") as f:
    for k, lines in listing.items():
        if "dgemm_nt_kernel<128, false>" in k:
            f.write("==== %s (%d instructions)\n" % (k, len(lines)))
            f.write("\n".join(lines) + "\n")
print("wrote", prefix + "_summary.txt", prefix + "_dgemm_nt_kernel.txt")
