"""Synthetic PRGs and reads of the shapes BASELINE.json names.

PRGs are emitted directly as the integer string `gramtools build` would store in gram_dir/prg
(1..4 = ACGT, odd >= 5 site entry, even = allele separator / site end;
gramtools/commands/build/vcf_to_prg_string.py:81-101 for the `--vcf` route). Reads are error-free
substrings of random haplotype paths, on a random strand.
"""
import numpy as np


def master_seeds(seed, n):
    """Per-read selection seeds: the first n raw outputs of std::mt19937(seed)
    (handle_read_file, quasimap.cpp:136-137). numpy's MT19937 seeded with the same 32-bit integer via
    `init_genrand` produces the identical stream."""
    bg = np.random.MT19937()
    st = bg.state
    key = np.zeros(624, dtype=np.uint32)
    x = np.uint64(seed & 0xFFFFFFFF)
    key[0] = x
    for i in range(1, 624):
        x = (np.uint64(1812433253) * (x ^ (x >> np.uint64(30))) + np.uint64(i)) & np.uint64(0xFFFFFFFF)
        key[i] = x
    st["state"]["key"] = key
    st["state"]["pos"] = 624
    bg.state = st
    return bg.random_raw(n).astype(np.uint32)


def make_snp_prg(ref_len, n_sites, seed, min_spacing=2):
    """Random reference + biallelic SNPs -> (prg uint32, ref uint8 codes 1..4, site_pos, alt)."""
    rng = np.random.default_rng(seed)
    ref = rng.integers(1, 5, size=ref_len, dtype=np.uint8)
    # positions without replacement, not adjacent (adjacency is exercised by the nested generator)
    cand = np.sort(rng.choice(ref_len - 2, size=min(n_sites * 2 + 16, ref_len - 2), replace=False)) + 1
    keep = [cand[0]]
    for p in cand[1:]:
        if p - keep[-1] >= min_spacing:
            keep.append(p)
    pos = np.asarray(keep, dtype=np.int64)
    if pos.size > n_sites:
        pos = np.sort(rng.choice(pos, size=n_sites, replace=False))
    n_sites = pos.size
    alt = ((ref[pos].astype(np.int64) - 1 + rng.integers(1, 4, size=n_sites)) % 4 + 1).astype(np.uint8)
    out = np.zeros(ref_len + 4 * n_sites, dtype=np.uint32)
    is_site = np.zeros(ref_len, dtype=np.int64)
    is_site[pos] = 1
    before = np.cumsum(is_site) - is_site          # sites strictly before i
    out_pos = np.arange(ref_len) + 4 * before + is_site  # ref base of a site sits after the odd marker
    out[out_pos] = ref
    sp = out_pos[pos]
    ids = 5 + 2 * np.arange(n_sites, dtype=np.uint32)
    out[sp - 1] = ids
    out[sp + 1] = ids + 1
    out[sp + 2] = alt
    out[sp + 3] = ids + 1
    return out, ref, pos, alt


def make_indel_prg(ref_len, n_sites, seed, snp_frac=0.8):
    """SNP + indel sites as config 4/5 describe them: 80 % SNPs, 10 % deletions (REF of
    2-10 bases, ALT = its first base), 10 % insertions (ALT = REF base + 1-9 random bases), emitted as
    `odd REF even ALT even` (vcf_to_prg_string.py:81-101). Plain loop: meant for test-sized PRGs."""
    rng = np.random.default_rng(seed)
    ref = rng.integers(1, 5, size=ref_len)
    starts = np.sort(rng.choice(np.arange(1, ref_len - 12, 14), size=min(n_sites, (ref_len - 13) // 14), replace=False))
    out, prev, sid = [], 0, 5
    for p in starts:
        p = int(p)
        out += [int(x) for x in ref[prev:p]]
        r = rng.random()
        if r < snp_frac:
            refa = [int(ref[p])]
            alta = [int((ref[p] - 1 + rng.integers(1, 4)) % 4 + 1)]
        elif r < snp_frac + (1 - snp_frac) / 2:
            n = int(rng.integers(2, 11))
            refa = [int(x) for x in ref[p:p + n]]
            alta = [int(ref[p])]
        else:
            refa = [int(ref[p])]
            alta = [int(ref[p])] + [int(x) for x in rng.integers(1, 5, size=int(rng.integers(1, 10)))]
        out += [sid] + refa + [sid + 1] + alta + [sid + 1]
        sid += 2
        prev = p + len(refa)
    out += [int(x) for x in ref[prev:]]
    return np.asarray(out, dtype=np.uint32)


def make_indel_prg_np(ref_len, n_sites, seed, snp_frac=0.8):
    """Vectorised make_indel_prg for config-4/5-sized PRGs: 80 % SNPs, 10 % deletions (REF of 2-10
    bases, ALT = its first base), 10 % insertions (ALT = REF base + 1-9 random bases), emitted as
    `odd REF even ALT even`. Returns (prg uint32, ref uint8 codes, sites) with sites = dict of arrays
    start / ref_len / alt_off / alt_len / alt_bases for `indel_haplotypes`."""
    rng = np.random.default_rng(seed)
    ref = rng.integers(1, 5, size=ref_len, dtype=np.uint8)
    stride = 13  # REF alleles are at most 10 bases: sites never touch
    n_sites = min(n_sites, (ref_len - 14) // stride)
    slots = np.sort(rng.choice((ref_len - 14) // stride, size=n_sites, replace=False))
    start = (slots * stride + 1 + rng.integers(0, 2, size=n_sites)).astype(np.int64)
    r = rng.random(n_sites)
    is_del = (r >= snp_frac) & (r < snp_frac + (1 - snp_frac) / 2)
    is_ins = r >= snp_frac + (1 - snp_frac) / 2
    rlen = np.ones(n_sites, dtype=np.int64)
    rlen[is_del] = rng.integers(2, 11, size=int(is_del.sum()))
    alen = np.ones(n_sites, dtype=np.int64)
    alen[is_ins] = 1 + rng.integers(1, 10, size=int(is_ins.sum()))
    alt_off = np.zeros(n_sites + 1, dtype=np.int64)
    np.cumsum(alen, out=alt_off[1:])
    alt_bases = rng.integers(1, 5, size=int(alt_off[-1]), dtype=np.uint8)
    first = alt_off[:-1]
    # first ALT base: SNP = a different base; deletion / insertion = the REF allele's first base
    snp = ~(is_del | is_ins)
    alt_bases[first[snp]] = ((ref[start[snp]].astype(np.int64) - 1 + rng.integers(1, 4, size=int(snp.sum()))) % 4 + 1)
    alt_bases[first[~snp]] = ref[start[~snp]]
    # output index of every reference base: +1 at a site start (odd marker), +(2 + alen) after the REF allele
    shift = np.zeros(ref_len + 1, dtype=np.int64)
    np.add.at(shift, start, 1)
    np.add.at(shift, start + rlen, 2 + alen)
    shift = np.cumsum(shift)[:ref_len]
    out = np.zeros(ref_len + int((3 + alen).sum()), dtype=np.uint32)
    out[np.arange(ref_len) + shift] = ref
    ids = 5 + 2 * np.arange(n_sites, dtype=np.uint32)
    o_start = start + shift[start]                 # first REF-allele base in the PRG
    out[o_start - 1] = ids
    sep = o_start + rlen                           # even marker after the REF allele
    out[sep] = ids + 1
    within = np.arange(int(alt_off[-1])) - np.repeat(first, alen)
    out[np.repeat(sep + 1, alen) + within] = alt_bases
    out[sep + 1 + alen] = ids + 1
    sites = dict(start=start, ref_len=rlen, alt_off=alt_off, alt_len=alen, alt_bases=alt_bases)
    return out, ref, sites


def indel_haplotypes(ref, sites, n_hap, seed):
    """Random haplotypes of a make_indel_prg_np PRG: every site takes REF or ALT with equal probability."""
    rng = np.random.default_rng(seed)
    start, rlen, alen, alt_off, alt_bases = (sites[k] for k in ("start", "ref_len", "alt_len", "alt_off", "alt_bases"))
    haps = []
    for _ in range(n_hap):
        pick = rng.integers(0, 2, size=start.size).astype(bool)
        keep = np.ones(ref.size + 1, dtype=np.int64)
        # drop the REF allele of ALT-carrying sites, then splice the ALT bases in front of what follows it
        d = np.zeros(ref.size + 1, dtype=np.int64)
        np.add.at(d, start[pick], 1)
        np.add.at(d, start[pick] + rlen[pick], -1)
        dropped = np.cumsum(d)[:ref.size] > 0
        ins_at = start[pick] + rlen[pick]          # ALT bases go before reference position ins_at
        ins_len = alen[pick]
        add = np.zeros(ref.size + 1, dtype=np.int64)
        np.add.at(add, ins_at, ins_len)
        kept = ~dropped
        new_pos = np.cumsum(kept.astype(np.int64) + add[:ref.size]) - kept  # index of base i if kept (after inserts at i)
        total = int(kept.sum() + ins_len.sum())
        h = np.zeros(total, dtype=np.uint8)
        h[new_pos[kept]] = ref[kept]
        # inserted runs end right before new_pos[ins_at] (or at the end of the haplotype)
        end = np.where(ins_at < ref.size, new_pos[np.minimum(ins_at, ref.size - 1)] - 0, total)
        # when the base at ins_at is itself dropped (cannot happen: sites never touch) new_pos would be off
        within = np.arange(int(ins_len.sum())) - np.repeat(np.cumsum(ins_len) - ins_len, ins_len)
        src = np.repeat(alt_off[:-1][pick], ins_len) + within
        h[np.repeat(end - ins_len, ins_len) + within] = alt_bases[src]
        del keep
        haps.append(h)
    return haps


def snp_haplotypes(ref, pos, alt, n_hap, seed):
    rng = np.random.default_rng(seed)
    haps = []
    for _ in range(n_hap):
        h = ref.copy()
        pick = rng.integers(0, 2, size=pos.size).astype(bool)
        h[pos[pick]] = alt[pick]
        haps.append(h)
    return haps


def make_nested_prg(n_loci, locus_len, seed, max_depth=3, spacer=200, distinct=False):
    """Bracket-grammar PRG with nesting, empty alleles (direct deletions) and adjacent sites.
    distinct=True: the alleles of a site are pairwise different, as in a PRG built from an MSA (make_prg collapses
    identical sequences); with the default, short alleles can coincide, and every such site doubles the search
    states of the reads crossing it (kept for the tests: it is the worst case for the general machinery)."""
    rng = np.random.default_rng(seed)
    out = []
    next_id = [5]

    def seq(n):
        return [int(x) for x in rng.integers(1, 5, size=n)]

    def site(depth, budget):
        sid = next_id[0]
        next_id[0] += 2
        res = [sid]
        n_all = int(rng.integers(2, 5))
        empty_used = False
        seen = []
        for a in range(n_all):
            if a:
                res.append(sid + 1)
            r = rng.random()
            if r < 0.12 and not empty_used and a > 0:
                empty_used = True  # empty allele = direct deletion
                continue
            al = body(depth + 1, max(1, int(budget * rng.uniform(0.2, 0.6))))
            while distinct and al in seen:
                al = body(depth + 1, max(2, int(budget * rng.uniform(0.2, 0.6))))
            seen.append(al)
            res += al
        res.append(sid + 1)
        return res

    def body(depth, budget):
        res = []
        if depth <= max_depth and budget >= 6 and rng.random() < 0.5:
            if rng.random() < 0.7:
                res += seq(int(rng.integers(1, max(2, budget // 3))))
            res += site(depth, budget // 2)
            if rng.random() < 0.25 and depth <= max_depth:  # adjacent sites `][`
                res += site(depth, budget // 3)
            if rng.random() < 0.7:
                res += seq(int(rng.integers(1, max(2, budget // 3))))
        else:
            res += seq(int(rng.integers(1, max(2, min(budget, 12)))))
        return res

    out += seq(spacer)
    for _ in range(n_loci):
        remaining = locus_len
        while remaining > 0:
            step = int(rng.integers(8, 40))
            out += seq(step)
            out += site(1, 40)
            remaining -= step + 40
        out += seq(spacer)
    return np.asarray(out, dtype=np.uint32)


def random_haplotype(prg, rng):
    """One random path through any (nested) PRG: uniform allele choice at every site."""
    prg = np.asarray(prg)
    n = prg.size
    # number of alleles per site and matching positions
    n_alleles = {}
    last = {}
    for i, m in enumerate(prg):
        m = int(m)
        if m > 4 and m % 2 == 0:
            n_alleles[m - 1] = n_alleles.get(m - 1, 0) + 1
            last[m] = i
    out = []
    # stack of [site, chosen allele, current allele]; `skip` depth counter for unchosen alleles
    stack = []
    active = True
    act_stack = []
    for i in range(n):
        m = int(prg[i])
        if m <= 4:
            if active:
                out.append(m)
        elif m % 2 == 1:
            act_stack.append(active)
            chosen = int(rng.integers(0, n_alleles[m])) if active else -1
            stack.append([m, chosen, 0])
            active = active and chosen == 0
        else:
            s = stack[-1]
            if i == last[m]:
                stack.pop()
                active = act_stack.pop()
            else:
                s[2] += 1
                active = act_stack[-1] and s[1] == s[2]
    return np.asarray(out, dtype=np.uint8)


def sample_reads(haps, n_reads, read_len, seed, frac_garbage=0.0, frac_n=0.0):
    """Error-free reads from random haplotypes, random strand. Returns (bases uint8 concatenated,
    offsets uint64). `frac_garbage` adds uniformly random reads, `frac_n` empty reads (non-ACGT)."""
    rng = np.random.default_rng(seed)
    hap_idx = rng.integers(0, len(haps), size=n_reads)
    L = read_len
    reads = np.zeros((n_reads, L), dtype=np.uint8)
    for h, hap in enumerate(haps):
        sel = np.nonzero(hap_idx == h)[0]
        if sel.size == 0:
            continue
        if hap.size < L:
            raise ValueError("haplotype shorter than the read length")
        starts = rng.integers(0, hap.size - L + 1, size=sel.size)
        idx = starts[:, None] + np.arange(L)[None, :]
        reads[sel] = hap[idx]
    rc = rng.random(n_reads) < 0.5
    reads[rc] = (5 - reads[rc])[:, ::-1]
    if frac_garbage > 0:
        g = rng.random(n_reads) < frac_garbage
        reads[g] = rng.integers(1, 5, size=(int(g.sum()), L), dtype=np.uint8)
    lens = np.full(n_reads, L, dtype=np.uint64)
    if frac_n > 0:
        lens[rng.random(n_reads) < frac_n] = 0
    if (lens == L).all():
        bases = reads.reshape(-1)
    else:
        bases = np.concatenate([reads[i, :int(lens[i])] for i in range(n_reads)]) if n_reads else np.zeros(0, np.uint8)
    offsets = np.zeros(n_reads + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    return np.ascontiguousarray(bases), offsets
