"""One .naf from several record-aligned shards: the multi-GPU form of ennaf (include/nafgpu.h).

Every rank owns a record-aligned piece of the text.  The protocol has ONE small exchange in the middle and one
gather at the end:

    begin    each rank parses / splits / packs its shard                     (shard encoder: ``begin``)
    link     all-gather of 9 integers per rank; every rank derives, on its own, what it must know about its
             neighbours: the global index of its first base (4-bit nibble parity), the case run that crosses
             into it, whether it is last                                      (``link_for``)
    finish   nibble shift, boundary mask runs, zstd blocks                    (shard encoder: ``finish``)
    gather   rank 0 lays out header + per stream {VLE sizes, frame header, the blocks of rank 0, 1, ...} and the
             ranks send their blocks straight into place                      (``encode_sharded``)

The reference has no counterpart (it is single-threaded); the format facts relied on are SURVEY A.1 / A.2: one zstd
frame per stream (unnaf's sequence / quality loops stop after the first frame), any block size <= 128 KB, and the
concatenation rules of the six streams (ids / comments / lengths / quality simply concatenate, the 4-bit stream
and the mask run lengths are global).

Everything here is plain integer bookkeeping plus ``torch.distributed`` calls on uint8 tensors, so it runs over NCCL
with CUDA tensors (the product) and over gloo with CPU tensors (tests, with a CPU shard encoder built on the oracle).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

from . import container

STREAMS = ("ids", "comments", "lengths", "mask", "sequence", "quality")
N_COUNTS = 9


@dataclass
class Counts:
    n_records: int = 0
    n_bases: int = 0
    longest_line: int = 0
    n_flips: int = 0
    last_flip: int = 0
    first_code: int = 0
    first_case: int = 0
    last_case: int = 0
    format: int = 0          # 1 FASTA, 2 FASTQ, 0 empty shard

    def as_list(self) -> List[int]:
        return [self.n_records, self.n_bases, self.longest_line, self.n_flips, self.last_flip, self.first_code,
                self.first_case, self.last_case, self.format]

    @classmethod
    def from_list(cls, v: Sequence[int]) -> "Counts":
        return cls(*[int(x) for x in v])


@dataclass
class Link:
    bases_before: int
    run_carry: int
    prev_last_case: int
    next_first_code: int
    is_last: int
    store_qual: int = 0          # some shard is FASTQ: every shard contributes (possibly empty) blocks to the quality stream


def link_for(all_counts: Sequence[Counts], rank: int) -> Link:
    """What shard `rank` needs to know about the others (mirrors nafgpu_shard_link)."""
    bases_before, last_flip, prev_case = 0, None, 0
    for c in all_counts[:rank]:
        if c.n_bases:
            if c.first_case != prev_case:        # a case change sits exactly on that shard's first base
                last_flip = bases_before
            if c.n_flips:
                last_flip = bases_before + c.last_flip
            prev_case = c.last_case
        bases_before += c.n_bases
    nxt = [c for c in all_counts[rank + 1:] if c.n_bases]
    return Link(bases_before=bases_before, run_carry=bases_before - (last_flip or 0), prev_last_case=prev_case,
                next_first_code=nxt[0].first_code if nxt else 0, is_last=int(rank == len(all_counts) - 1),
                store_qual=int(any(c.format == 2 for c in all_counts)))


def container_layout(seq_type: int, title: Optional[bytes], line_length: Optional[int], all_counts: Sequence[Counts],
                     raw: Sequence[Sequence[int]], body: Sequence[Sequence[int]], store_mask: bool, store_qual: bool,
                     window_log: int = 17):
    """-> (total size, [(offset, bytes)] small host-written pieces, [[offset of rank r's blocks of stream k]])

    ennaf.c:538-589: magic, version, [type], flags, separator, VLE line length, VLE N, [title], then per stream
    VLE(original size) VLE(compressed size - 4) and the frame without its 4-byte magic: FHD 0x00, window byte, blocks.
    For the sequence stream "original size" is the number of bases (ennaf.c:582)."""
    world = len(all_counts)
    n_records = sum(c.n_records for c in all_counts)
    longest = max([c.longest_line for c in all_counts] + [0])
    head = bytearray(b"\x01\xf9\xec")
    head += bytes([1]) if seq_type == 0 else bytes([2, seq_type])
    present = [True, True, True, store_mask, True, store_qual]
    flags = (0x40 if title is not None else 0) | 0x20 | 0x10 | 0x08 | (0x04 if store_mask else 0) | 0x02 | (0x01 if store_qual else 0)
    head += bytes([flags, 0x20])
    head += container.put_vle(line_length if line_length is not None else longest)
    head += container.put_vle(n_records)
    if title is not None:
        head += container.put_vle(len(title)) + title
    pieces, at, pos = [], [[0] * 6 for _ in range(world)], 0
    pending = bytes(head)
    for k in range(6):
        if not present[k]:
            continue
        orig = sum(c.n_bases for c in all_counts) if k == 4 and seq_type < 2 else sum(raw[r][k] for r in range(world))
        comp = 2 + sum(body[r][k] for r in range(world))                 # FHD + window byte + blocks
        pending += container.put_vle(orig) + container.put_vle(comp) + bytes([0x00, (window_log - 10) << 3])
        pieces.append((pos, pending))
        pos += len(pending)
        pending = b""
        for r in range(world):
            at[r][k] = pos
            pos += body[r][k]
    if pending:
        pieces.append((pos, pending))
        pos += len(pending)
    return pos, pieces, at


class GpuShardEncoder:
    """The product: libnafgpu.so (nafgpu_shard_begin / _finish / _fetch) on this rank's GPU."""

    def __init__(self, ctx, device_text: bool = False):
        self.ctx, self.device_text = ctx, device_text

    def begin(self, text, opts) -> Counts:
        from . import api
        c, self.info = self.ctx.shard_begin(text, opts, on_device=self.device_text)
        return Counts(c.n_records, c.n_bases, c.longest_line, c.n_flips, c.last_flip, c.first_code, c.first_case, c.last_case, c.format)

    def finish(self, link: Link):
        from . import api
        l = api.ShardLink()
        l.bases_before, l.run_carry, l.prev_last_case = link.bases_before, link.run_carry, link.prev_last_case
        l.next_first_code, l.is_last, l.store_qual = link.next_first_code, link.is_last, link.store_qual
        return self.ctx.shard_finish(l)

    def fetch(self, stream: int, dst):
        """dst: uint8 tensor (CUDA or CPU) of exactly the body size"""
        if dst.numel():
            self.ctx.shard_fetch(stream, dst.data_ptr())


def encode_sharded(encoder, text, opts, *, seq_type: int = 0, title: Optional[bytes] = None, line_length: Optional[int] = None,
                   group=None, device=None):
    """Collective over `group` (torch.distributed): every rank passes its record-aligned shard of the text; rank 0
    returns the .naf as a uint8 tensor on `device`, the other ranks return None."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    device = device or torch.device("cpu")
    counts = encoder.begin(text, opts)
    mine = torch.tensor(counts.as_list(), dtype=torch.int64, device=device)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    all_counts = [Counts.from_list(g.tolist()) for g in gathered]
    formats = {c.format for c in all_counts if c.format}
    if len(formats) > 1:
        raise ValueError("shards disagree about the input format (FASTA / FASTQ)")
    raw, body = encoder.finish(link_for(all_counts, rank))
    sizes = torch.tensor(list(raw) + list(body), dtype=torch.int64, device=device)
    all_sizes = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    raws = [s.tolist()[:6] for s in all_sizes]
    bodies = [s.tolist()[6:] for s in all_sizes]
    store_qual = 2 in formats
    store_mask = seq_type < 2 and not bool(getattr(opts, "no_mask", 0))
    total, pieces, at = container_layout(seq_type, title, line_length, all_counts, raws, bodies, store_mask, store_qual)
    present = [True, True, True, store_mask, True, store_qual]
    # the gather: every rank's blocks go straight to their place in rank 0's image of the file
    if rank == 0:
        out = torch.empty(total, dtype=torch.uint8, device=device)
        for pos, data in pieces:
            out[pos:pos + len(data)] = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(device)
        reqs = []
        for k in range(6):
            if not present[k]:
                continue
            encoder.fetch(k, out[at[0][k]:at[0][k] + bodies[0][k]])
            for r in range(1, world):
                if bodies[r][k]:
                    reqs.append(dist.irecv(out[at[r][k]:at[r][k] + bodies[r][k]], src=dist.get_global_rank(group, r) if group is not None else r, group=group, tag=k))
        for q in reqs:
            q.wait()
        return out
    for k in range(6):
        if not present[k] or not bodies[rank][k]:
            continue
        buf = torch.empty(bodies[rank][k], dtype=torch.uint8, device=device)
        encoder.fetch(k, buf)
        dist.send(buf, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group, tag=k)
    return None


def encode_shards_local(encoders, texts, opts, *, seq_type: int = 0, title: Optional[bytes] = None, line_length: Optional[int] = None) -> bytes:
    """The same protocol without a process group: one encoder object per shard, all in this process (tests; a
    single GPU working through a file in pieces)."""
    import torch
    all_counts = [e.begin(t, opts) for e, t in zip(encoders, texts)]
    sized = [e.finish(link_for(all_counts, r)) for r, e in enumerate(encoders)]
    raws, bodies = [list(s[0]) for s in sized], [list(s[1]) for s in sized]
    formats = {c.format for c in all_counts if c.format}
    store_qual = 2 in formats
    store_mask = seq_type < 2 and not bool(getattr(opts, "no_mask", 0))
    total, pieces, at = container_layout(seq_type, title, line_length, all_counts, raws, bodies, store_mask, store_qual)
    out = torch.empty(total, dtype=torch.uint8)
    for pos, data in pieces:
        out[pos:pos + len(data)] = torch.frombuffer(bytearray(data), dtype=torch.uint8)
    present = [True, True, True, store_mask, True, store_qual]
    for r, e in enumerate(encoders):
        for k in range(6):
            if present[k] and bodies[r][k]:
                e.fetch(k, out[at[r][k]:at[r][k] + bodies[r][k]])
    return out.numpy().tobytes()


def split_records(text: bytes, pieces: int) -> List[bytes]:
    """Record-aligned pieces of a FASTA text ('>' at a line start is unambiguous) or a 4-line-per-record FASTQ text
    (record starts are found by exact line counting from the top: a quality line may begin with '@')."""
    import numpy as np
    n = len(text)
    if pieces <= 1 or n == 0:
        return [text]
    a = np.frombuffer(text, dtype=np.uint8)
    first = text.lstrip()[:1]
    cuts = [0]
    if first == b">":
        for k in range(1, pieces):
            at = text.find(b"\n>", k * n // pieces)
            cuts.append(n if at < 0 else max(at + 1, cuts[-1]))
    else:
        nl = np.flatnonzero(a == 10)                   # exact: line index of every newline
        for k in range(1, pieces):
            target = k * n // pieces
            j = int(np.searchsorted(nl, target))       # first newline at or after target
            j += (-(j + 1)) % 4                        # advance to the newline that ends a 4-line record
            cuts.append(n if j >= len(nl) else max(int(nl[j]) + 1, cuts[-1]))
    cuts.append(n)
    return [text[cuts[i]:cuts[i + 1]] for i in range(pieces)]


def split_records_gpu(ctx, text, pieces: int, on_device: bool = False):
    """The same cut points found on the GPU (nafgpu_record_cuts: newline ordinals by a prefix sum over tiles) -- what the
    multi-GPU tools use; `text` may be a (device address, nbytes) pair.  Returns the list of pieces + 1 offsets."""
    return ctx.record_cuts(text, pieces, on_device=on_device)


def record_range(n_records: int, rank: int, world: int):
    """records [first, first + count) of rank `rank` when `n_records` are dealt out evenly"""
    first = n_records * rank // world
    return first, n_records * (rank + 1) // world - first


RANGE_VIEWS = ("default", "fasta", "fastq", "sequences", "ids", "names")   # nafgpu_dec_opts.first_record / n_records apply


def decode_shard(ctx, naf, rank: int, world: int, view: str = "default", **kw) -> bytes:
    """This rank's piece of the text of a .naf file: the ranks' pieces, concatenated in rank order, are the full output
    (nafgpu_dec_opts.first_record / n_records; sequence and quality blocks outside the range are not decoded when the
    frames carry no sequences, as ours do).  Every rank needs the whole file (it is the small side: broadcast it)."""
    if view not in RANGE_VIEWS:
        # --seq / --4bit / --lengths / --mask / --charcount are not per-record outputs: the library ignores a record range
        # for them, so one rank (piece) produces the whole output and the others contribute nothing
        return ctx.decode(naf, view, **kw) if rank == 0 else b""
    n = container.read_header(naf if isinstance(naf, bytes) else bytes(naf)).n_sequences
    first, count = record_range(n, rank, world)
    if count == 0:
        return b""
    return ctx.decode(naf, view, first_record=first, n_records=count, **kw)


# ---------------------------------------------------------------------------------------------------------------------
# One GPU working through a file that does not fit in HBM: the shard protocol run sequentially.
#
# The reference streams its input through 16 KB / 128 KB windows into one temp file per stream and concatenates them at
# the end (ennaf.c:538-589, compressor.c:150).  Here a "window" is a record-aligned piece of a few hundred MB: piece i is
# begun on one context while piece i-1 -- which needs piece i's first base code for its last nibble -- is finished on
# the other; the zstd blocks of each stream accumulate on the host (``sink``) and the container is laid out at the end.

def iter_record_pieces(chunks, piece_bytes: int):
    """Record-aligned pieces of roughly `piece_bytes` from an iterable of byte chunks (a file read in order).
    FASTA: cut before a '>' at a line start.  FASTQ: cut after every 4th line, counted from the top of the file (a
    quality line may begin with '@'), i.e. 4-line records without blank lines -- what split_records accepts too."""
    import numpy as np
    buf, fmt, lines = bytearray(), None, 0              # lines: newlines in the pieces already yielded (FASTQ)
    for chunk in chunks:
        buf += chunk
        if fmt is None:
            head = bytes(buf[:4096]).lstrip()
            if not head:
                continue
            fmt = ">" if head[:1] == b">" else "@"
        while len(buf) >= piece_bytes:
            if fmt == ">":
                at = buf.rfind(b"\n>", 0, len(buf))
                cut = at + 1 if at >= 0 else 0
            else:
                nl = np.flatnonzero(np.frombuffer(buf, dtype=np.uint8) == 10)
                ends = nl[(lines + np.arange(1, len(nl) + 1)) % 4 == 0]       # newlines that end a record
                cut = int(ends[-1]) + 1 if len(ends) else 0
                if cut:
                    lines += int(np.searchsorted(nl, cut))
            if cut == 0:
                break                                   # one record longer than a piece: keep reading
            yield bytes(buf[:cut])
            del buf[:cut]
    if buf or fmt is None:
        yield bytes(buf)


class MemorySink:
    """Per-stream accumulation of zstd blocks in host memory (the reference uses one temp file per stream)."""

    def __init__(self):
        self.parts = [[] for _ in range(6)]

    def buffer(self, k: int, n: int):
        import torch
        t = torch.empty(n, dtype=torch.uint8)
        self.parts[k].append(t)
        return t

    def stream(self, k: int) -> bytes:
        return b"".join(p.numpy().tobytes() for p in self.parts[k])


def encode_stream(encoders, pieces, opts, *, seq_type: int = 0, title: Optional[bytes] = None, line_length: Optional[int] = None,
                  sink=None) -> bytes:
    """Sequential form of encode_sharded for ONE device: `encoders` are two shard encoders (two contexts on the same GPU)
    used alternately, `pieces` an iterable of record-aligned texts (iter_record_pieces).  At most two pieces are resident
    at any time.  Returns the .naf."""
    sink = sink or MemorySink()
    all_counts: List[Counts] = []
    raws: List[List[int]] = []
    bodies: List[List[int]] = []
    pending = None                                      # (encoder, index into all_counts) begun, not finished
    carry = b""
    turn = 0

    def finish(enc, idx, is_last):
        link = link_for(all_counts, idx)
        link.is_last = int(is_last)
        raw, body = enc.finish(link)
        raws.append(list(raw)); bodies.append(list(body))
        for k in range(6):
            if body[k]:
                enc.fetch(k, sink.buffer(k, body[k]))

    for piece in pieces:
        piece = carry + piece
        carry = b""
        enc = encoders[turn]
        c = enc.begin(piece, opts)
        if pending is not None and c.n_bases == 0 and (sum(x.n_bases for x in all_counts) & 1):
            # the piece before this one ends on a low nibble and this piece has no base to complete it with: take the
            # next piece into this one and begin again (begin on the same context drops the abandoned shard)
            carry = piece
            continue
        formats = {x.format for x in all_counts + [c] if x.format}
        if len(formats) > 1:
            raise ValueError("pieces disagree about the input format (FASTA / FASTQ)")
        all_counts.append(c)
        if pending is not None:
            finish(pending[0], pending[1], False)
        pending = (enc, len(all_counts) - 1)
        turn ^= 1
    if carry:                                           # trailing pieces without bases: one last shard
        enc = encoders[turn]
        all_counts.append(enc.begin(carry, opts))
        if pending is not None:
            finish(pending[0], pending[1], False)
        pending = (enc, len(all_counts) - 1)
    if pending is None:                                 # no input at all: one empty shard
        all_counts.append(encoders[0].begin(b"", opts))
        pending = (encoders[0], 0)
    finish(pending[0], pending[1], True)

    formats = {c.format for c in all_counts if c.format}
    store_qual = 2 in formats
    store_mask = seq_type < 2 and not bool(getattr(opts, "no_mask", 0))
    total, pieces_hdr, at = container_layout(seq_type, title, line_length, all_counts, raws, bodies, store_mask, store_qual)
    out = bytearray(total)
    for pos, data in pieces_hdr:
        out[pos:pos + len(data)] = data
    present = [True, True, True, store_mask, True, store_qual]
    for k in range(6):
        if present[k]:
            s = sink.stream(k)
            assert len(s) == sum(b[k] for b in bodies)
            out[at[0][k]:at[0][k] + len(s)] = s         # the shards' blocks of a stream are contiguous, in order
    return bytes(out)


def decode_stream(ctx, naf, write, pieces: int, view: str = "default", **kw) -> int:
    """unnaf for outputs larger than HBM: the records in `pieces` consecutive ranges, each decoded by one call and handed to
    `write(bytes)` in order.  Returns the number of bytes written."""
    n = 0
    for r in range(max(1, pieces)):
        part = decode_shard(ctx, naf, r, max(1, pieces), view, **kw)
        write(part)
        n += len(part)
    return n
