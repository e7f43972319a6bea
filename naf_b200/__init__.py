"""naf_b200 — B200 (sm_100a) implementation of the NAF encode/decode hot path.

Thin ctypes binding over ``libnafgpu.so`` (C ABI in ``include/nafgpu.h``).  The host-side names
mirror the reference's two tools: :func:`ennaf` = what ``ennaf``'s ``main()`` does after the command
line is parsed (``ennaf/src/ennaf.c:433``), :func:`unnaf` = ``unnaf``'s ``main()``
(``unnaf/src/unnaf.c:356``).  There is no CPU path: importing works without a GPU (so the symbol
table can be checked), but creating a context raises :class:`NafGpuError` if no sm_100 device exists
or the extension has not been built.
"""
from .api import (  # noqa: F401
    NafGpu,
    NafGpuError,
    EncOpts,
    DecOpts,
    EncInfo,
    Timing,
    ShardCounts,
    ShardLink,
    ennaf,
    unnaf,
    load_library,
    library_path,
    DNA, RNA, PROTEIN, TEXT,
    OUT_DEFAULT, OUT_FASTA, OUT_FASTQ, OUT_SEQ, OUT_SEQUENCES, OUT_4BIT, OUT_IDS, OUT_NAMES, OUT_LENGTHS,
    OUT_MASK, OUT_CHARCOUNT,
)

__version__ = "0.1.0"
