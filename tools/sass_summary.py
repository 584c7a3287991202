"""Counts the Blackwell-specific SASS mnemonics per kernel of the built library (cuobjdump -sass): the evidence that the hot
kernels are tcgen05 / TMEM / TMA code.  python tools/sass_summary.py > profiles/rNN_sass_mnemonics.txt"""
import collections
import hashlib
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "wavedm_b200", "libwavedm_b200.so")
PAT = re.compile(r"\b(UTCHMMA|UTMALDG|UTMASTG|LDTM|UTCBAR|UTCATOMSWS|SYNCS|HMMA|LDSM|LDGSTS|ACQBULK|UTMAPF|UBLKCP)\b")


def main():
    sha = hashlib.sha256(open(LIB, "rb").read()).hexdigest()[:16]
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, cnt = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            cnt[cur] = collections.Counter()
        elif cur:
            mm = PAT.search(line)
            if mm:
                cnt[cur][mm.group(1)] += 1
    names = subprocess.run(["cu++filt"] + list(cnt.keys()), capture_output=True, text=True).stdout.splitlines()
    tot, lines = collections.Counter(), []
    for n, c in zip(names, cnt.values()):
        if not c:
            continue
        tot.update(c)
        n = re.sub(r"void |\(anonymous namespace\)::|wdm::", "", re.sub(r"\(.*", "", n))
        lines.append(f"{n[:72]:72s} " + " ".join(f"{a}={b}" for a, b in sorted(c.items())))
    print(f"# SASS mnemonics per kernel of wavedm_b200/libwavedm_b200.so (sha256[:16] {sha}), from `cuobjdump -sass`:")
    print("# UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld (TMEM -> registers), "
          "UTCBAR = tcgen05.commit,")
    print("# UTCATOMSWS = TMEM alloc / dealloc, SYNCS = mbarrier ops, HMMA = mma.sync (HFRM small-C kernels), LDSM = ldmatrix, "
          "LDGSTS = cp.async")
    print("# totals: " + " ".join(f"{a}={b}" for a, b in sorted(tot.items())))
    print("\n".join(sorted(lines)))


if __name__ == "__main__":
    main()
