"""profiles/r2_sass_grep.txt: per-kernel counts of the SASS mnemonics that prove which hardware paths a kernel uses."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "carla_ppo_b200", "libcarla_ppo_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTMALDG[.\w]*|UTMAPF[.\w]*|UTMASTG[.\w]*|UBLKCP[.\w]*|UTCHMMA[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|LDTM|STTM|LDGSTS|RED\.[.\w]*|FFMA|MEMBAR[.\w]*|SYNCS[.\w]*)")
counts, order, cur = collections.defaultdict(collections.Counter), [], None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::|cpb::|\(.*$", "", cur)
        order.append(cur)
        continue
    m = pat.search(ln)
    if m and cur:
        k = m.group(1)
        k = "SYNCS" if k.startswith("SYNCS") else ("FFMA" if k == "FFMA" else k)
        counts[cur][k] += 1
out = ["# SASS mnemonic counts of libcarla_ppo_b200.so (cuobjdump -sass, sm_100a), per kernel; built from the sources at this commit",
       "# UTMALDG = cp.async.bulk.tensor (TMA tensor-map load), UTMAPF = its L2-prefetch form, UBLKCP = cp.async.bulk (no tensor map),",
       "# UTCHMMA = tcgen05.mma (kind::tf32; .2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / tcgen05.st, UTCBAR = tcgen05.commit,",
       "# UTCATOMSWS = tcgen05.alloc / dealloc, LDGSTS = cp.async, RED = red.global.add (bias-gradient column sums), SYNCS = mbarrier ops", ""]
for k in order:
    c = counts[k]
    if not any(x.startswith(("UT", "LDTM", "STTM", "UBLKCP", "LDGSTS")) for x in c) and "edge" not in k and "ppo" not in k:
        continue
    out.append("%-60s %s" % (k[:60], " ".join("%s=%d" % kv for kv in sorted(c.items()))))
open(os.path.join(ROOT, "profiles", "r2_sass_grep.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:5] + [l for l in out if "tc2_tapgemm_kernel<128, true, false>" in l or "tc3" in l]))
