"""SASS evidence of the Blackwell-native kernels: static counts of the tcgen05 / TMEM / TMA / cluster / peer-memory mnemonics per
kernel of libmixq_b200.so, plus the full listing of the dominant decode kernel (mixq_gemm_dequant_fat_kernel).
    python profiles/sass_evidence.py  ->  profiles/r2_sass_evidence.txt, profiles/r2_sass_fat_kernel.sass"""
import collections
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "mixq_tensorrt_llm_b200" / "libmixq_b200.so"
KEYS = re.compile(r"\b(UTCIMMA|UTCHMMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMASTG|UTMACCTL|UBLKCP|STAS|SYNCS|UCGABAR_ARV|UCGABAR_WAIT|ACQBULK|"
                  r"REDG|ATOMG|MEMBAR|LDGSTS|I2FP|F2FP|MUFU|HFMA2|REDUX|CREDUX|MULTIMEM|ERRBAR)[A-Z0-9_.]*")
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
cur, counts, listing, names = None, collections.OrderedDict(), [], {}
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    if "fat_kernel" in cur and not re.fullmatch(r"\s*/\* 0x[0-9a-f]+ \*/\s*", ln):   # drop the second encoding word of every instruction
        listing.append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln))
    for k in KEYS.finditer(ln.split("/*")[1] if ln.strip().startswith("/*") and ln.count("/*") > 1 else ln):
        counts[cur][k.group(0)] += 1
dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
out = ["SASS mnemonics (cuobjdump -sass libmixq_b200.so), static instruction counts per kernel.",
       "UTCIMMA / UTCHMMA = tcgen05.mma kind::i8 / kind::f16 (.2CTA = cta_group::2); LDTM / STTM = tcgen05.ld / tcgen05.st (TMEM);",
       "UTMALDG / UTMASTG = TMA tensor loads / stores; UBLKCP = cp.async.bulk; STAS = st.async into another CTA's shared memory;",
       "UTCBAR = tcgen05.commit; UCGABAR = cluster barrier; REDG.*.SYS / MEMBAR.ALL.SYS = system-scope signalling over NVLink peer memory.", ""]
for mangled, name in zip(counts, dem):
    c = counts[mangled]
    if not any(k.startswith(("UTC", "LDTM", "UTMA", "STAS", "REDG", "LDGSTS", "HFMA2")) for k in c):
        continue
    short = re.sub(r"mixq::\(anonymous namespace\)::", "", name)
    short = re.sub(r"\(CUtensorMap_st.*", "", short)[:160]
    out.append(short)
    out.append("   " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))
    out.append("")
(ROOT / "profiles" / "r2_sass_evidence.txt").write_text("\n".join(out))
(ROOT / "profiles" / "r2_sass_fat_kernel.sass").write_text("\n".join(listing) + "\n")
print(len(counts), "kernels;", len(listing), "lines of the fat kernel")
