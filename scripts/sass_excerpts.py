"""profiles/r02_sass_*.txt: SASS evidence of the round-2 kernels, from the objects `python -m goliath_b200.build` leaves
under goliath_b200/build/ (cuobjdump -sass; no GPU needed).  Per object: every kernel with its instruction count and
the mnemonics that prove (or disprove) the Blackwell paths — UTC*MMA / LDTM / UTMALDG / UBLKCP / SYNCS / REDG / MUFU —
plus the full listing of the named hot kernel."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "goliath_b200", "build")
WATCH = ("FFMA2", "UTCHMMA", "UTCQMMA", "UTCMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "REDG", "RED.", "ATOMS", "ATOMG",
         "MUFU", "SHFL", "HMMA", "LDGSTS", "BAR.SYNC", "MATCH", "VOTE")
TARGETS = {
    "splat_blend_mom": ("blend_fwd_ilp_kernelILi4ELb1ELb0", "profiles/r02_sass_splat_blend_mom.txt"),
    "deconv_wnub": ("deconv4x4s2_bwd_data_wide_kernel", "profiles/r02_sass_deconv_wnub.txt"),
    "deconv_tc": ("deconv_tc_kernel", "profiles/r02_sass_deconv_tc.txt"),
    "mvp_raymarch": ("raymarch_fwd_kernelILb0ELb0", "profiles/r02_sass_mvp_raymarch.txt"),
    "splat_bin_tiles": ("rank_sort_coop_kernelILi8", "profiles/r02_sass_splat_bin_tiles.txt"),
}


def main():
    for obj, (hot, out) in TARGETS.items():
        path = os.path.join(BUILD, obj + ".o")
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        funcs, cur = collections.OrderedDict(), None
        for line in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = m.group(1)
                funcs[cur] = []
            elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
                funcs[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line).rstrip())
        with open(os.path.join(ROOT, out), "w") as f:
            f.write("# cuobjdump -sass goliath_b200/build/%s.o  (sm_100a; nvcc 12.9, flags of goliath_b200/build.py)\n" % obj)
            f.write("# per kernel: instructions, then counts of the mnemonics that matter\n")
            for name, ins in funcs.items():
                short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:150]
                cnt = collections.Counter()
                for l in ins:
                    for w in WATCH:
                        if w in l and not (w == "HMMA" and "UTCHMMA" in l):
                            cnt[w.rstrip(".")] += 1
                f.write("%6d  %s\n        %s\n" % (len(ins), short, " ".join("%s=%d" % kv for kv in sorted(cnt.items()))))
            hit = [n for n in funcs if hot in n]
            if hit:
                f.write("\n# ---- full listing: %s\n" % subprocess.run(["c++filt", hit[0]], capture_output=True, text=True).stdout.strip()[:200])
                body = funcs[hit[0]]
                f.write("\n".join(body[:900]) + "\n")
                if len(body) > 900:
                    f.write("# ... %d more instructions (regenerate with scripts/sass_excerpts.py for the full listing)\n" % (len(body) - 900))
        print(out, sum(len(v) for v in funcs.values()), "instructions in", len(funcs), "kernels")


if __name__ == "__main__":
    sys.exit(main())
