"""TEST INFRASTRUCTURE: writes a Dazzler read database (.db stub, .idx, .bps) and a .las file in the
layouts falcon_b200/csrc/fcx_dazz.cu reads (DAZZ_DB DB.h: HITS_DB / HITS_READ; DALIGNER align.h:
Overlap I/O records), and the LA4Falcon -fo text of the same overlaps, so that the binary path can be
checked against the text path.  The real Dazzler tools are not available here."""
import os
import struct

CODE = {65: 0, 67: 1, 71: 2, 84: 3}
COMP = bytes.maketrans(b"ACGT", b"TGCA")
DB_BEST = 0x800


def revcomp(s: bytes) -> bytes:
    return s.translate(COMP)[::-1]


def write_db(dirname, root, reads, cutoff=0, all_flag=1, flags=None):
    """reads: list of bytes (ACGT).  flags: per-read flag words (default DB_BEST)."""
    bps = bytearray()
    recs = []
    for i, r in enumerate(reads):
        boff = len(bps)
        n = len(r)
        for j in range(0, n, 4):
            b = 0
            for k in range(4):
                b <<= 2
                if j + k < n:
                    b |= CODE[r[j + k]]
            bps.append(b)
        fl = DB_BEST if flags is None else flags[i]
        recs.append(struct.pack("<iii4xqqi4x", i, n, 0, boff, -1, fl))
    tot = sum(len(r) for r in reads)
    maxlen = max((len(r) for r in reads), default=0)
    hdr = struct.pack("<iiii4fi4xqiiiii4xQi4xQQQ", len(reads), len(reads), cutoff, all_flag, .25, .25, .25, .25,
                      maxlen, tot, len(reads), 0, 0, 0, 0, 0, 0, 0, 0, 0)
    assert len(hdr) == 112 and all(len(x) == 40 for x in recs)
    with open(os.path.join(dirname, "." + root + ".idx"), "wb") as f:
        f.write(hdr + b"".join(recs))
    with open(os.path.join(dirname, "." + root + ".bps"), "wb") as f:
        f.write(bytes(bps))
    with open(os.path.join(dirname, root + ".db"), "w") as f:
        f.write("files =         1\n%10d %s %s\nblocks =         1\nsize =       200 cutoff = %9d all = %d\n         0         0\n%10d%10d\n"
                % (len(reads), root, root, cutoff, all_flag, len(reads), len(reads)))
    return os.path.join(dirname, root + ".db")


def write_las(path, overlaps, tspace=100):
    """overlaps: list of dicts aread, bread, comp, abpos, aepos, bbpos, bepos (file order)."""
    with open(path, "wb") as f:
        f.write(struct.pack("<qi", len(overlaps), tspace))
        for o in overlaps:
            tlen = 2 * max(1, (o["aepos"] - o["abpos"]) // tspace)
            f.write(struct.pack("<6iI2i4x", tlen, 0, o["abpos"], o["bbpos"], o["aepos"], o["bepos"],
                                1 if o["comp"] else 0, o["aread"], o["bread"]))
            f.write(bytes(tlen * (1 if tspace <= 125 else 2)))


def la4falcon_text(reads, overlaps, seed_cutoff=0):
    """What `LA4Falcon -H<seed_cutoff> -fo` prints for these overlaps (reads: trimmed-DB order)."""
    out = []
    p_aread = -1
    for o in overlaps:
        a, b = o["aread"], o["bread"]
        alen, blen = len(reads[a]), len(reads[b])
        if o["abpos"] != 0 and o["bbpos"] != 0:
            continue
        if o["aepos"] != alen and o["bepos"] != blen:
            continue
        if alen < seed_cutoff:
            continue
        if a != p_aread:
            if p_aread != -1:
                out.append(b"+ +\n")
            out.append(b"%08d %s\n" % (a, reads[a]))
            p_aread = a
        out.append(b"%08d %s\n" % (b, revcomp(reads[b]) if o["comp"] else reads[b]))
    if p_aread != -1:
        out.append(b"+ +\n")
    out.append(b"- -\n")
    return b"".join(out)
