"""Per-kernel counts of the SASS mnemonics that prove the data path of lib/libgla_cuda.so:
TMA = UTMALDG, tcgen05 = UTCHMMA / UTCBAR / UTCATOMSWS / LDTM, FP64 tensor pipe = DMMA, TF32 mma.sync = HMMA,
cp.async = LDGSTS, mbarrier = SYNCS, st.async to a peer CTA = STAS, cluster barrier = UCGABAR.   usage: python tools/sass_excerpt.py > profiles/...txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "genericlinearalgebra.jl_b200", "lib", "libgla_cuda.so")
PAT = re.compile(r"\b(UTCHMMA|UTCBAR|UTCATOMSWS|LDTM|UTMALDG|UTMACCTL|DMMA|HMMA|LDGSTS|SYNCS|STAS|UCGABAR_ARV|UCGABAR_WAIT|DFMA|FFMA)(\.[A-Z0-9_.]+)?")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = {}
per = collections.OrderedDict()
fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        per[fn] = collections.Counter()
        continue
    if fn is None:
        continue
    m = PAT.search(line)
    if m:
        key = m.group(1) if m.group(1) in ("DFMA", "FFMA") else m.group(1) + (m.group(2) or "")
        per[fn][key] += 1
dem = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass genericlinearalgebra.jl_b200/lib/libgla_cuda.so: per-kernel counts of the mnemonics that prove the data path")
print("# (tcgen05 = UTCHMMA / UTCBAR / UTCATOMSWS / LDTM, TMA = UTMALDG, FP64 tensor pipe = DMMA.8x8x4, TF32 mma.sync = HMMA.1688,")
print("#  cp.async = LDGSTS, mbarrier = SYNCS, st.async into a peer CTA of the cluster = STAS, cluster barrier = UCGABAR); kernels without any of them are omitted\n")
tot = collections.Counter()
for (fn, c), d in zip(per.items(), dem):
    tc = {k: v for k, v in c.items() if k not in ("DFMA", "FFMA")}
    tot.update(tc)
    if not tc:
        continue
    print(d[:150])
    print("    " + ", ".join(f"{k} x{v}" for k, v in sorted(tc.items())) + f"   (DFMA x{c['DFMA']}, FFMA x{c['FFMA']})")
print("\nlibrary totals: " + ", ".join(f"{k} x{v}" for k, v in sorted(tot.items())))
