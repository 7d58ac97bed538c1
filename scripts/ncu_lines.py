#!/usr/bin/env python
"""Per-source-line instruction counts for one kernel of an .ncu-rep.

  python scripts/ncu_lines.py REP KERNEL_SUBSTRING [TOP_N] [LIB.so] [MANGLED_SUBSTRING]

KERNEL_SUBSTRING selects the launch in the report (demangled name); MANGLED_SUBSTRING (e.g. ILj255E for the <255>
instantiation; default: KERNEL_SUBSTRING) selects the ONE function section of the cubin whose line table is used — with
several instantiations of a template in the library a loose match joins the counts with the wrong function's lines.
An instruction is attributed to the innermost frame of its inline chain that lies in this repository's sources (CUDA's
own headers — shuffle and atomic wrappers — are skipped), so intrinsics count for the line that calls them.

ncu's SASS page gives executed warp instructions / thread instructions / stall samples per instruction;
nvdisasm --print-line-info-inline on the cubin of the SAME build gives file:line per instruction (needs
-lineinfo).  The two are joined on the instruction's offset within the kernel.
"""
import csv, os, re, subprocess, sys, tempfile
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
mangled = sys.argv[5] if len(sys.argv) > 5 else kern
lib = sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "microsimulation_b200", "libmsim_b200.so")

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
inst, cur, hdr = {}, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        cur = r[1] if kern in r[1] else None
        hdr = None
        continue
    if r[0] == "Address":
        hdr = r
        continue
    if cur is None or hdr is None or cur in inst and False:
        continue
    d = dict(zip(hdr, r))
    inst.setdefault(cur, []).append(d)
if not inst:
    sys.exit("kernel not found in report")
name, recs = next(iter(inst.items()))
base = int(recs[0]["Address"], 16)

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info-inline", cubin], capture_output=True, text=True).stdout
OURS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "microsimulation_b200", "csrc")
lines, infn, where, in_chain, matched = {}, False, None, False, []
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
    if m:
        infn = mangled in m.group(1)
        if infn:
            matched.append(m.group(1))
        where, in_chain = None, False
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        # a run of //## lines is one inline chain, innermost location first: take the innermost frame in our sources
        here = (m.group(1).split("/")[-1], int(m.group(2)))
        ours = os.path.exists(os.path.join(OURS, here[0]))
        if not in_chain:
            where, where_ours = here, ours
        elif not where_ours and ours:
            where, where_ours = here, True
        in_chain = True
        continue
    in_chain = False
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines[int(m.group(1), 16)] = where

if len(set(matched)) != 1:
    sys.exit("the mangled substring %r matches %d function sections: %s" % (mangled, len(set(matched)), sorted(set(matched))[:6]))
agg, tot, tots = {}, 0.0, 0.0
for d in recs:
    off = int(d["Address"], 16) - base
    ie = float(d.get("Instructions Executed") or 0)
    te = float(d.get("Thread Instructions Executed") or 0)
    sm = float(d.get("# Samples") or 0)
    w = lines.get(off) or ("?", 0)
    a = agg.setdefault(w, [0.0, 0.0, 0.0])
    a[0] += ie; a[1] += te; a[2] += sm
    tot += ie; tots += sm
print("%s\nwarp instructions executed %.0f, stall samples %.0f" % (name, tot, tots))
files = {}
for (f, l), (ie, te, sm) in agg.items():
    a = files.setdefault(f, [0.0, 0.0, 0.0]); a[0] += ie; a[1] += te; a[2] += sm
print("\nby file:   inst%   samples%   lanes")
for f, (ie, te, sm) in sorted(files.items(), key=lambda kv: -kv[1][0]):
    print("%-24s %5.1f%%  %5.1f%%  %5.1f" % (f, 100 * ie / tot, 100 * sm / max(tots, 1), te / ie if ie else 0))
src_cache = {}
def src(f, l):
    for d in ("microsimulation_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][l - 1].strip()[:100] if 0 < l <= len(src_cache[p]) else ""
    return ""
print("\nby line:   inst%   samples%   lanes")
for (f, l), (ie, te, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%%  %5.1f%%  %5.1f  %s:%d  %s" % (100 * ie / tot, 100 * sm / max(tots, 1), te / ie if ie else 0, f, l, src(f, l)))
