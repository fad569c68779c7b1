"""SASS evidence for profiles/: instruction mix and the Blackwell-specific instructions of the two kernels of the bench
job, from `cuobjdump -sass` of the built objects (no GPU needed)."""
import collections
import re
import subprocess
import sys

def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, name = [], None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, cur
            name, cur = m.group(1), []
        elif name:
            cur.append(line)
    if name:
        yield name, cur

def demangle(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()

def report(obj, want, keep):
    for name, lines in functions(obj):
        d = demangle(name)
        if not re.search(want, d):
            continue
        ops = []
        for l in lines:
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if m:
                ops.append(m.group(1))
        mix = collections.Counter(o.split(".")[0] for o in ops)
        print(f"### `{d.replace('nk::(anonymous namespace)::', '')}`\n")
        print(f"{len(ops)} SASS instructions (static).  Mix: " + ", ".join(f"{k} {v}" for k, v in mix.most_common(16)) + "\n")
        print("```")
        seen = collections.Counter()
        for l in lines:
            m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
            if not m:
                continue
            for k in keep:
                if re.search(k, m.group(2)) and seen[k] < 3:
                    seen[k] += 1
                    print(f"/*{m.group(1)}*/ {m.group(2).strip()}")
        print("```\n")

if __name__ == "__main__":
    print("# r02 — SASS excerpts (`cuobjdump -sass`, sm_100a, built by `python -m neurokmer_b200.build`)\n")
    print("Single `sm_100a` image per object.  What to look for: `UBLKCP` = 1-D TMA bulk copy (`cp.async.bulk`), `SYNCS` = mbarrier "
          "arrive/try_wait, `REDG.E.ADD` = fire-and-forget L2 reduction (`red.global.add.u32`), `IMAD.X` = the high word of a 64-bit add on "
          "the FMA pipe, `DFMA`/`DADD`/`DMUL` = the exact modulo on the FP64 pipe, `SHF.L.W` = one half of a 64-bit rotate.  There is no "
          "dense contraction anywhere on this path, so no `UTC*MMA` / `LDTM` (tcgen05 / TMEM) and no `UTMALDG` (tensor-map TMA): the tiles "
          "are linear byte ranges.\n")
    report("neurokmer_b200/build/nk_count.o", r"count_kernel<true, 5, 3, true, false>", [r"UBLKCP", r"SYNCS", r"REDG", r"IMAD\.X", r"DFMA", r"SHF\.L\.W", r"ATOMG"])
    report("neurokmer_b200/build/nk_post.o", r"post_kernel", [r"LDG\.E\.STRONG\.SYS|LD\.E\.STRONG\.SYS|\.SYS", r"ATOMS", r"ATOMG|REDG", r"BAR\.SYNC", r"MEMBAR|ERRBAR", r"CCTL"])
    report("neurokmer_b200/build/nk_parse.o", r"fa_write_kernel", [r"VOTE|BALLOT", r"STS", r"STG\.E\.128", r"POPC"])
