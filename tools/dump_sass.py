"""Per-kernel SASS listings of the shipped library -> profiles/sass_r2/ (north_star: "a committed SASS listing per kernel").

    python tools/dump_sass.py

One file per kernel: the instruction stream as cuobjdump prints it (encodings stripped), headed by its mnemonic histogram.
Template families (conv_fwd2_tc_kernel has 22 instantiations) are written once per representative instantiation; the index
lists every kernel of the library with its instruction count and the tcgen05 / TMEM / TMA mnemonics it contains."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "adaptivepnp_sci_b200", "libsci_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass_r2")
KEY = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMACMDFLUSH", "SYNCS", "ELECT", "HMMA", "FFMA", "DFMA", "RED", "ATOM")
# instantiations written in full for the template families: <EPI_WG, MODE, HALF>
FULL = ("conv_fwd2_tc_kernelILi2ELi0ELb1E", "conv_fwd2_tc_kernelILi2ELi1ELb1E", "conv_fwd2_tc_kernelILi2ELi2ELb1E",
        "conv_fwd2_tc_kernelILi1ELi0ELb0E", "conv_fwd2_tc_kernelILi2ELi1ELb0E", "conv_fwd_tc_kernelILb1E", "conv_fwd_tc_kernelILb0E",
        "project_kernel_vecILi8ELi4E", "project_kernel_vecILi8ELi1E")


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except OSError:
        return n


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    os.makedirs(OUT, exist_ok=True)
    for f in os.listdir(OUT):
        os.remove(os.path.join(OUT, f))
    kernels, cur = [], None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = [m.group(1), []]
            kernels.append(cur)
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4})\*/\s+(.*?);", line)
        if m and cur is not None:
            cur[1].append("/*%s*/  %s ;" % (m.group(1), m.group(2).strip()))
    index = ["# SASS of libsci_b200.so (sm_100a), `cuobjdump -sass`; one row per kernel: instructions, tensor-core / TMEM / TMA / barrier mnemonics\n"]
    for name, ins in kernels:
        hist = collections.Counter()
        for i in ins:
            op = i.split("*/", 1)[1].split()
            op = [o for o in op if not o.startswith("@")][0].rstrip(";")
            hist[op.split(".")[0]] += 1
        short = re.sub(r"^_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+", "", name)
        short = re.sub(r"^_Z\d*", "", short)
        keys = ", ".join("%s %d" % (k, hist[k]) for k in KEY if hist[k])
        dem = demangle(name).replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
        index.append("%-70s %6d instr   %s" % (dem[:70], len(ins), keys))
        family = any(t in name for t in ("conv_fwd2_tc_kernelI", "conv_fwd_tc_kernelI", "project_kernel_vecI"))
        if family and not any(t in name for t in FULL):
            continue
        fn = re.sub(r"[^A-Za-z0-9_]", "_", dem)[:80] + ".sass"
        with open(os.path.join(OUT, fn), "w") as f:
            f.write("// %s\n// %d instructions; histogram: %s\n" % (demangle(name), len(ins), ", ".join("%s %d" % kv for kv in hist.most_common(14))))
            f.write("\n".join(ins) + "\n")
    with open(os.path.join(OUT, "INDEX.txt"), "w") as f:
        f.write("\n".join(index) + "\n")
    print("wrote %d listings + INDEX.txt to %s" % (len(os.listdir(OUT)) - 1, OUT))


main()
