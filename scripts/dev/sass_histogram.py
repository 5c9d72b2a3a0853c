"""Per-kernel SASS opcode evidence for the Blackwell-native paths (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st ->
LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP; legacy mma.sync would show as HMMA). Runs here (no GPU needed):
    python scripts/dev/sass_histogram.py > profiles/r02_sass_histogram.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "MUFU", "REDG", "ATOMG"]


def histogram(lib: Path):
    out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(1)
            kernels[cur]["total"] += 1
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    kernels[cur][o] += 1
    return kernels


def demangle(names):
    res = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return res.stdout.splitlines() if res.returncode == 0 else names


def main():
    for libname in ("libb200pt.so", "libb200pt_fp16.so"):
        lib = ROOT / "multimodal_llm_pretraining_b200" / libname
        if not lib.exists():
            continue
        ks = histogram(lib)
        names = demangle(list(ks))
        print(f"# {libname}: {len(ks)} kernels; SASS opcode counts per kernel (cuobjdump -sass, CUDA 12.9, sm_100a)")
        print(f"# {'kernel':100s} " + " ".join(f"{o:>8s}" for o in ["total"] + OPS))
        tot = collections.Counter()
        for (k, c), nm in zip(ks.items(), names):
            nm = re.sub(r"\(.*", "", nm).replace("void ", "")
            print(f"{nm[:100]:102s} " + " ".join(f"{c.get(o, 0):8d}" for o in ["total"] + OPS))
            tot.update(c)
        print(f"{'TOTAL':102s} " + " ".join(f"{tot.get(o, 0):8d}" for o in ["total"] + OPS))
        print()


if __name__ == "__main__":
    sys.exit(main())
