#!/usr/bin/env python
"""Extract the Metropolis loop of the timed sweep kernel -- k_sweep2<2,FAST> (two customers per thread; `python
tools/sass_mhloop.py r02 1` for the one-customer k_sweep<2,FAST>) -- from the built library (cuobjdump -sass; no GPU needed),
classify its instructions and write profiles/<tag>_sweep_mhloop_sass.md"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = os.path.join(ROOT, "mcmc_clv_model_b200", "libclv_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cpt = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pat = r"k_sweep2ILi2ELi0E" if cpt == 2 else r"k_sweepILi2ELi0ELb0"
m = re.search(r"Function : (\S*" + pat + r"\S*)(.*?)(?=\n\s*Function : |\Z)", sass, re.S)
name, text = m.group(1), m.group(2)
ins = [(int(a, 16), t.strip()) for a, t in re.findall(r"/\*([0-9a-f]{4})\*/\s+(.*?);", text)]
cos = [a for a, t in ins if "MUFU.COS" in t]
back = []
for a, t in ins:
    mm = re.search(r"BRA\S* (?:\S+, )?0x([0-9a-f]+)", t)
    if mm and int(mm.group(1), 16) < a:
        back.append((a, int(mm.group(1), 16)))
# the Metropolis loop: the innermost backward branch whose body holds 2 x cpt MUFU.COS (two t3 variates per customer and step)
a, b = min(((a, b) for a, b in back if sum(1 for c in cos if b <= c <= a) == 2 * cpt), key=lambda t: t[0] - t[1])
body = [(x, t) for x, t in ins if b <= x <= a]
# the rarely executed regions: every forward branch of the loop body that jumps over a CALL skips one (the exact fp64
# re-decision of the accept tie zone with its clip to +-70: ~1e-4 of the steps; in the CLV_E32=0 / one-customer builds also
# the separate clip calls)
def opcode(t):
    return (t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]
rare = set()
calls = [x for x, t in body if "CALL" in t]
for x, t in body:
    mm = re.search(r"BRA\S* (?:\S+, )?0x([0-9a-f]+)", t)
    if mm:
        tgt = int(mm.group(1), 16)
        if tgt > x and any(x < cx < tgt for cx in calls):
            rare |= {y for y, _ in body if x < y < tgt}
hot = [(x, t) for x, t in body if x not in rare]
groups = collections.OrderedDict([
    ("Philox4x32-10 (IMAD.WIDE + LOP3 + PRMT)", lambda t: opcode(t) in ("LOP3", "PRMT") or "IMAD.WIDE" in t),
    ("fp64 (DFMA / DMUL / DADD / DSETP)", lambda t: opcode(t) in ("DFMA", "DMUL", "DADD", "DSETP")),
    ("fp32 + SFU (FFMA / FMUL / FADD / FSETP / MUFU / conversions)", lambda t: opcode(t) in ("FFMA", "FMUL", "FADD", "FSETP", "MUFU", "I2FP", "F2F", "I2F")),
    ("selects / predicates (FSEL / SEL / PLOP3 / ISETP / VIMNMX)", lambda t: opcode(t) in ("FSEL", "SEL", "PLOP3", "ISETP", "VIMNMX")),
    ("constant / uniform loads (LDC / LDCU)", lambda t: opcode(t) in ("LDC", "LDCU")),
    ("shared / local memory (LDS / STS / LDL / STL + address)", lambda t: opcode(t) in ("LDS", "STS", "LDL", "STL", "IADD3")),
    ("moves / integer housekeeping / branches", lambda t: True)])
cnt = collections.Counter()
for x, t in hot:
    for g, f in groups.items():
        if f(t):
            cnt[g] += 1
            break
out = os.path.join(ROOT, "profiles", f"{tag}_sweep_mhloop_sass.md" if cpt == 2 else f"{tag}_sweep1_mhloop_sass.md")
with open(out, "w") as f:
    f.write(f"# {tag}: SASS of the Metropolis loop of `{'k_sweep2' if cpt == 2 else 'k_sweep'}<2,FAST>` (cuobjdump -sass of the committed build)\n\n")
    f.write(f"`{name}`: loop 0x{b:04x} .. 0x{a:04x}, {len(body)} instructions in the loop body ({cpt} customer(s) per thread and trip), of which "
            f"{len(body) - len(hot)} sit in the rarely taken regions (the exact fp64 re-decision of the accept tie zone, which also clips a "
            f"proposal beyond +-70: ~1e-4 of the steps), leaving {len(hot)} instructions per trip = **{len(hot) / cpt:.0f} instructions per "
            f"customer Metropolis step** on the hot path:\n\n| group | instructions per trip |\n|---|---|\n")
    for g in groups:
        f.write(f"| {g} | {cnt[g]} |\n")
    nloc = sum(1 for x, t in body if 'LDL' in t or 'STL' in t)
    f.write(f"\nLocal-memory instructions in the loop: {nloc} ({'no spill' if nloc == 0 else 'register spills at the 64 registers of 8 blocks per SM'}).\n\n```\n")
    for x, t in body:
        f.write(f"{x:04x}{' r' if x in rare else '  '} {t}\n")
    f.write("```\n(`r` marks the rarely executed regions.)\n")
print("wrote", out, "hot path:", len(hot), dict(cnt))
