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
a, b = min(((a, b) for a, b in back if b < cos[0] < a), key=lambda t: t[0] - t[1])
body = [(x, t) for x, t in ins if b <= x <= a]
# the rarely executed regions: the clip calls and the exact exp() of the accept tie zone (between the DSETP that follows
# the fp32 screen and the BSYNC that closes it)
def opcode(t):
    return (t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]
rare = set()
# every tie zone: from the DSETP.GE (d >= 0) that follows a failed fp32 screen to the DSETP.GT (exp(d) > u) that closes it
pos = 0
while True:
    tie0 = next((x for x, t in body if x > pos and "DSETP.GE" in t), None)
    tie1 = next((x for x, t in body if tie0 and x > tie0 and t.startswith("DSETP.GT")), None)
    if not (tie0 and tie1):
        break
    rare |= {x for x, t in body if tie0 < x < tie1}
    pos = tie1
# every clip: from the branch that skips it to the last CALL of the group
calls = [x for x, t in body if "CALL" in t]
for cx in calls:
    lo = max((x for x, t in body if x < cx and "BRA" in t and x not in rare), default=cx)
    rare |= {x for x, t in body if lo < x <= cx + 0x30}
hot = [(x, t) for x, t in body if x not in rare]
groups = collections.OrderedDict([
    ("Philox4x32-10 (IMAD.WIDE + LOP3 + PRMT)", lambda t: opcode(t) in ("LOP3", "PRMT") or "IMAD.WIDE" in t),
    ("fp64 (DFMA / DMUL / DADD / DSETP)", lambda t: opcode(t) in ("DFMA", "DMUL", "DADD", "DSETP")),
    ("fp32 + SFU (FFMA / FMUL / FADD / FSETP / MUFU / conversions)", lambda t: opcode(t) in ("FFMA", "FMUL", "FADD", "FSETP", "MUFU", "I2FP", "F2F", "I2F")),
    ("selects / predicates (FSEL / SEL / PLOP3 / ISETP / VIMNMX)", lambda t: opcode(t) in ("FSEL", "SEL", "PLOP3", "ISETP", "VIMNMX")),
    ("constant / uniform loads (LDC / LDCU)", lambda t: opcode(t) in ("LDC", "LDCU")),
    ("shared-memory table (LDS + address)", lambda t: opcode(t) in ("LDS", "IADD3")),
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
            f"{len(body) - len(hot)} sit in the rarely taken regions (the clip of a proposal beyond +-70: 1 step in 50 000; the exact fp64 "
            f"`exp` of the accept tie zone: ~1e-5 of the steps), leaving {len(hot)} instructions per trip = **{len(hot) / cpt:.0f} instructions per "
            f"customer Metropolis step** on the hot path:\n\n| group | instructions per trip |\n|---|---|\n")
    for g in groups:
        f.write(f"| {g} | {cnt[g]} |\n")
    f.write(f"\nLocal-memory instructions in the loop: {sum(1 for x, t in body if 'LDL' in t or 'STL' in t)} (no spill).\n\n```\n")
    for x, t in body:
        f.write(f"{x:04x}{' r' if x in rare else '  '} {t}\n")
    f.write("```\n(`r` marks the rarely executed regions.)\n")
print("wrote", out, "hot path:", len(hot), dict(cnt))
