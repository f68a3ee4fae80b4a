"""Static SASS counts of the k_build_trees builds in a library: instructions, local-memory (spill) loads / stores,
atomics, barriers, global loads / stores.  python tools/sass_stats.py drift_b200/libdrift_b200.so"""
import subprocess, sys, re
for f in sys.argv[1:]:
    out = subprocess.run(["cuobjdump", "-sass", f], capture_output=True, text=True).stdout
    parts = re.split(r"\n\s*Function : ", out)
    for p in parts[1:]:
        name = p.split("\n", 1)[0]
        if "k_build_trees" not in name: continue
        m = re.search(r"k_build_treesILi(\d)", name)
        lines = [l for l in p.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        c = lambda pat: sum(1 for l in lines if re.search(pat, l))
        print(f.split("/")[-1], m.group(1) if m else "old8", "instr", len(lines), "LDL", c(r"\bLDL"), "STL", c(r"\bSTL"), "ATOM", c(r"\bATOM|\bRED"), "ATOMS", c(r"\bATOMS"), "BAR", c(r"\bBAR"), "LDG", c(r"\bLDG"), "STG", c(r"\bSTG"))
