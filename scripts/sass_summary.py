"""SASS evidence per kernel of libcaffedistri_b200.so (run here, no GPU needed):
   python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, re, subprocess, sys
so = "caffeonspark_b200/libcaffedistri_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kernels, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = demangle(m.group(1))
        cur = re.sub(r"cosb::\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(cosb::SyncParams.*", "", cur).replace("void ", "")
        kernels[cur] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if m and cur:
        kernels[cur]["total"] += 1
        op = m.group(1)
        for key in ("LDGMC", "UBLKCP.S.G", "UBLKCP.G.S", "SYNCS", "MEMBAR.SC.SYS", "MEMBAR.ALL.SYS", "MEMBAR.ALL.GPU",
                    "FFMA", "FMUL", "FADD", "BAR.SYNC", "LDG.E.NA.128", "STG.E.128", "LDG.E.128.STRONG.SYS",
                    "STG.E.128.STRONG.SYS", "LDG.E.STRONG.SYS", "STG.E.STRONG.SYS", "LDG.E.64.STRONG.SYS",
                    "STG.E.64.STRONG.SYS", "REDG", "CCTL"):
            if op.startswith(key) and not (key == "STG.E.128" and "STRONG" in op) and not (key == "FADD" and False):
                kernels[cur][key] += 1
print("SASS evidence per kernel (cuobjdump -sass %s, sm_100a, nvcc 12.9)" % so)
print("""  LDG.E.NA.128 = ld.global.L1::no_allocate.v4.f32 (streaming 128-bit loads, local and peer)
  UBLKCP.S.G / UBLKCP.G.S = cp.async.bulk global->shared / shared->global (TMA); SYNCS.* = mbarrier
  LDGMC = multimem.ld_reduce (NVLS in-switch reduction); STG.E.128.STRONG.SYS in the nvls kernel = multimem.st
  LDG/STG.E.128.STRONG.SYS in the ll kernel = ld/st.relaxed.sys.v2.u64: two single-copy-atomic {payload, flag} words
  MEMBAR.ALL.SYS = fence.acq_rel.sys of the cross-GPU flag barrier (release before the flag store, acquire after
  the poll); the ll kernel has NONE (flag-in-data, fence-free)
  FFMA = 0 everywhere in the fused kernels: explicit __fmul_rn/__fadd_rn, nothing contracted -> bit-exact parity
""")
for k, c in kernels.items():
    print("%-58s %s" % (k[:58], "  ".join("%s=%d" % (a, b) for a, b in c.items())))
