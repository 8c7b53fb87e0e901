"""PLINK fileset helpers for the Python host layer (tests, bench, multi-GPU driver).

Mirrors the part of the reference's ``PioFile`` / ``PioFiles`` wrappers
(/root/reference/src/source.cc:187-361) that the hot path needs: open 1 or 22
bfiles, concatenate them into one global variant index, and map a global
variant index to (file, local row).  Text parsing follows libplinkio's
behaviour (fields split on runs of blanks, exactly 6 fields, integer
chromosome in an unsigned char: libplinkio/src/bim_parse.c:67-125, :224-226,
fam_parse.c:248-250).  The production host for the CLI is the C++ driver in
``simu_b200/host``; this module is its Python twin.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

BED_MAGIC = bytes([0x6C, 0x1B, 0x01])   # v1.00, SNP-major (bed_header.c:14-15, :31-46)


def row_bytes(n_samples: int) -> int:
    return (n_samples + 3) // 4           # bed_header.c:219-223


@dataclass
class Locus:
    chromosome: int
    name: str
    position: float
    bp_position: int
    allele1: str
    allele2: str


@dataclass
class Sample:
    fid: str
    iid: str
    father_iid: str
    mother_iid: str
    sex: str
    phenotype: str


class PlinkFormatError(RuntimeError):
    pass


def parse_bim(path: str) -> list[Locus]:
    out = []
    with open(path, "r") as f:
        for ln, line in enumerate(f, 1):
            tok = line.split()
            if not tok:
                continue
            if len(tok) != 6:
                raise PlinkFormatError(f"ERROR: Could not open {path[:-4]} (line {ln}: {len(tok)} fields, 6 expected)")
            try:
                chrom = int(tok[0], 10)
            except ValueError:
                raise PlinkFormatError(f"ERROR: Could not open {path[:-4]} (line {ln}: chromosome '{tok[0]}')")
            out.append(Locus(chrom & 0xFF, tok[1], float(tok[2]), int(tok[3]), tok[4], tok[5]))
    return out


def parse_fam(path: str) -> list[Sample]:
    out = []
    with open(path, "r") as f:
        for ln, line in enumerate(f, 1):
            tok = line.split()
            if not tok:
                continue
            if len(tok) != 6:
                raise PlinkFormatError(f"ERROR: Could not open {path[:-4]} (line {ln}: {len(tok)} fields, 6 expected)")
            out.append(Sample(*tok))
    return out


@dataclass
class PlinkFile:
    """One bfile prefix (PioFile, source.cc:187-267)."""
    prefix: str
    loci: list[Locus] = field(default_factory=list)
    samples: list[Sample] = field(default_factory=list)

    @classmethod
    def open(cls, prefix: str) -> "PlinkFile":
        for ext in (".fam", ".bim", ".bed"):
            if not os.path.exists(prefix + ext):
                raise PlinkFormatError(f"ERROR: Could not open {prefix}")
        pf = cls(prefix, parse_bim(prefix + ".bim"), parse_fam(prefix + ".fam"))
        with open(prefix + ".bed", "rb") as f:
            hdr = f.read(3)
        if len(hdr) < 3:
            raise PlinkFormatError(f"ERROR: Could not open {prefix}")
        if hdr[:2] != BED_MAGIC[:2] or hdr[2] != 1:
            raise PlinkFormatError(f"ERROR: Unsupported format in {prefix}, snps must be rows, samples must be columns")
        if pf.num_samples == 0:
            raise PlinkFormatError(f"ERROR: Unsupported format in {prefix}, rows appears to have zero size")
        return pf

    @property
    def num_samples(self) -> int:
        return len(self.samples)

    @property
    def num_loci(self) -> int:
        return len(self.loci)

    @property
    def bed_path(self) -> str:
        return self.prefix + ".bed"


class PlinkFiles:
    """Concatenation of bfiles in label order (PioFiles, source.cc:269-361)."""

    def __init__(self, prefixes: list[str]):
        self.files = [PlinkFile.open(p) for p in prefixes]
        n = {f.num_samples for f in self.files}
        if len(n) != 1:
            raise PlinkFormatError("ERROR: Inconsistent number of subjects in --bfile-chr input")   # source.cc:598
        self.num_samples = n.pop()
        self.offsets = np.cumsum([0] + [f.num_loci for f in self.files])
        self.num_variants = int(self.offsets[-1])

    @classmethod
    def from_bfile_chr(cls, pattern: str, labels=None) -> "PlinkFiles":
        labels = labels or [str(i) for i in range(1, 23)]          # source.cc:426-429
        if "@" not in pattern:
            pattern += "@"                                          # source.cc:435-436
        return cls([pattern.replace("@", lab) for lab in labels])

    def get_locus(self, variant_index: int) -> Locus:
        f = int(np.searchsorted(self.offsets, variant_index, side="right") - 1)
        return self.files[f].loci[variant_index - int(self.offsets[f])]

    def get_sample(self, i: int) -> Sample:
        return self.files[0].samples[i]                             # source.cc:309-311

    def split_by_file(self, variant_indices: np.ndarray):
        """ascending global indices -> list of (file_no, local_rows ndarray)."""
        v = np.asarray(variant_indices, dtype=np.int64)
        fno = np.searchsorted(self.offsets, v, side="right") - 1
        out = []
        for f in range(len(self.files)):
            sel = v[fno == f]
            if sel.size:
                out.append((f, sel - int(self.offsets[f])))
        return out

    def stage_causal_rows(self, ctx, variant_indices: np.ndarray) -> None:
        """Read the causal rows of every file into ctx's panel, in ascending
        variant order (the only rows find_freq / find_pheno ever read:
        source.cc:852-863, :969-980)."""
        v = np.asarray(variant_indices, dtype=np.int64)
        if v.size and np.any(np.diff(v) <= 0):
            raise ValueError("variant_indices must be strictly ascending")
        if ctx.panel_rows < max(1, v.size):
            from . import capi                   # compact causal panel, sample-major in groups of 16 rows (the CLI's layout)
            ctx.panel_alloc(max(1, v.size), capi.LAYOUT_SAMPLE16)
        dst = 0
        for f, local in self.split_by_file(v):
            ctx.panel_load_bed(self.files[f].bed_path, self.files[f].num_loci, local, dst_row=dst)
            dst += local.size


# ------------------------------------------------------------- GROUP4 layout (host restatement)

def interleave_group4(packed_rows: np.ndarray, n_samples: int, group_stride: int | None = None) -> np.ndarray:
    """Packed .bed rows [M, >=ceil(N/4)] -> the GROUP4 panel image [ceil(M/4), group_stride]:
    byte i of group g = the 2-bit fields of sample i for rows 4g..4g+3, row 4g in bits 0-1
    (include/simu_b200.h, SIMU_B200_LAYOUT_GROUP4).  What interleave_kernel (csrc/stage.cu)
    produces on the device; rows beyond M read as 00."""
    m = packed_rows.shape[0]
    if group_stride is None:
        group_stride = (n_samples + 127) // 128 * 128
    i = np.arange(n_samples)
    out = np.zeros(((m + 3) // 4, group_stride), dtype=np.uint8)
    for r in range(m):
        field = (packed_rows[r, i // 4] >> (2 * (i % 4))) & 3
        out[r // 4, :n_samples] |= (field << (2 * (r % 4))).astype(np.uint8)
    return out


def deinterleave_group4(panel: np.ndarray, n_rows: int, n_samples: int) -> np.ndarray:
    """The inverse: GROUP4 panel image -> packed .bed rows [n_rows, ceil(N/4)] (padding bits 00)."""
    B = row_bytes(n_samples)
    i = np.arange(n_samples)
    out = np.zeros((n_rows, B), dtype=np.uint8)
    for r in range(n_rows):
        field = (panel[r // 4, :n_samples] >> (2 * (r % 4))) & 3
        np.bitwise_or.at(out[r], i // 4, (field << (2 * (i % 4))).astype(np.uint8))
    return out


# ------------------------------------------------------------- SAMPLE16 layout (host restatement)

_BED_TO_DOSAGE = np.array([0, 3, 1, 2], np.uint32)     # .bed bit pairs 00, 01, 10, 11 -> dosage code (3 = missing)


def interleave_sample16(packed_rows: np.ndarray, n_samples: int, group_words: int | None = None) -> np.ndarray:
    """Packed .bed rows [M, >=ceil(N/4)] -> the SAMPLE16 panel image, uint32 [ceil(M/16), group_words]:
    word i of group g = the genotypes of sample i for rows 16g..16g+15, row 16g in bits 0-1, re-coded as the
    A2-allele dosage with 3 = missing (include/simu_b200.h, SIMU_B200_LAYOUT_SAMPLE16).  What
    interleave16_kernel (csrc/stage.cu) produces on the device; rows beyond M and samples >= N read as 0."""
    m = packed_rows.shape[0]
    if group_words is None:
        group_words = (n_samples + 127) // 128 * 128
    i = np.arange(n_samples)
    out = np.zeros(((m + 15) // 16, group_words), dtype=np.uint32)
    for r in range(m):
        field = (packed_rows[r, i // 4] >> (2 * (i % 4))) & 3
        out[r // 16, :n_samples] |= _BED_TO_DOSAGE[field] << np.uint32(2 * (r % 16))
    return out


def deinterleave_sample16(panel: np.ndarray, n_rows: int, n_samples: int) -> np.ndarray:
    """The inverse: SAMPLE16 panel image -> packed .bed rows [n_rows, ceil(N/4)] (padding bits 00)."""
    B = row_bytes(n_samples)
    i = np.arange(n_samples)
    to_bed = np.array([0, 2, 3, 1], np.uint8)
    out = np.zeros((n_rows, B), dtype=np.uint8)
    for r in range(n_rows):
        code = (panel[r // 16, :n_samples] >> np.uint32(2 * (r % 16))) & 3
        np.bitwise_or.at(out[r], i // 4, (to_bed[code] << (2 * (i % 4))).astype(np.uint8))
    return out


# ------------------------------------------------------------- fixture writers

def write_bed(path: str, packed_rows: np.ndarray, n_samples: int) -> None:
    """packed_rows: uint8 [M, >=row_bytes]; writes a v1.00 SNP-major .bed."""
    B = row_bytes(n_samples)
    with open(path, "wb") as f:
        f.write(BED_MAGIC)
        f.write(np.ascontiguousarray(packed_rows[:, :B]).tobytes())


def write_bim(path: str, n_loci: int, chromosome: int = 1, first_index: int = 0) -> None:
    """names rs<global index>, bp = local index + 1, alleles A/G."""
    with open(path, "w") as f:
        for i in range(n_loci):
            f.write(f"{chromosome}\trs{first_index + i}\t0\t{i + 1}\tA\tG\n")


def write_fam(path: str, n_samples: int) -> None:
    """README.md:58-60 ID style."""
    with open(path, "w") as f:
        for i in range(n_samples):
            f.write(f"id1_{i} id2_{i} 0 0 0 -9\n")


def write_fileset(prefix: str, packed_rows: np.ndarray, n_samples: int, chromosome: int = 1,
                  first_index: int = 0) -> None:
    write_bed(prefix + ".bed", packed_rows, n_samples)
    write_bim(prefix + ".bim", packed_rows.shape[0], chromosome, first_index)
    write_fam(prefix + ".fam", n_samples)
