"""profiles/r2_sass_excerpt.txt: per hot kernel of libb2vs.so, the counts of the Blackwell-only SASS
mnemonics (UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, UBLKCP = bulk copy, LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, ...) plus the first lines that carry them.
Run in the build container (cuobjdump needs no GPU)."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "cuvs-rag_b200", "libb2vs.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
want = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "UTCATOM", "SYNCS", "ELECT", "UTMAPF",
        "LDS", "STS", "ATOMG", "REDG", "RED"]
hot = ["bf_tc_kernel", "pq_tc_kernel", "merge_parts_kernel", "merge_splits_kernel", "ivf_flat_scan_kernel",
       "ivf_group_select_kernel", "ivf_seed_select_kernel", "gather_group", "pool_kernel", "unit_rows_kernel"]
elfs = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
out = [f"cuobjdump -sass {os.path.relpath(lib, ROOT)}  (embedded cubins: "
       f"{', '.join(sorted(set(re.findall(r'sm_[0-9a-z]+', elfs))))})", ""]
cur, per = None, collections.OrderedDict()
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        per[cur] = {"n": 0, "cnt": collections.Counter(), "first": {}}
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
    if cur and m:
        per[cur]["n"] += 1
        op = m.group(2).split(".")[0]
        if op in want:
            per[cur]["cnt"][op] += 1
            per[cur]["first"].setdefault(op, ln.strip()[:150])
for name, d in per.items():
    if not any(h in name for h in hot):
        continue
    short = re.sub(r"\(.*", "", name)
    out.append(f"== {short}   ({d['n']} SASS instructions)")
    out.append("   " + "  ".join(f"{k}={v}" for k, v in d["cnt"].most_common()))
    for op in ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR"):
        if op in d["first"]:
            out.append("     " + d["first"][op])
    out.append("")
path = os.path.join(ROOT, "profiles", "r2_sass_excerpt.txt")
open(path, "w").write("\n".join(out))
print("wrote", path, len(out), "lines")
