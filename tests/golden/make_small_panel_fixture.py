#!/usr/bin/env python
"""Decode the reference's own fixture data/small_panel.preprocessed.rda (the panel its testthat suite runs
buildGraph on, tests/testthat/test_functions.R:6-27) into tests/golden/small_panel_preprocessed.npz.

Runs only where the reference tree is mounted (this container); the .npz it writes is committed so that the GPU
box, which has no /root/reference, can run the same panel.  No R needed: the file is an xz-compressed RDX3 XDR
serialisation stream; the reader below implements the subset of R's unserialize (src/main/serialize.c, format
version 3) that the two Pagoda2 R6 objects use (environments, closures with byte code, S4 dgCMatrix, ALTREP
wrappers and compact sequences).

usage: python tests/golden/make_small_panel_fixture.py [path/to/small_panel.preprocessed.rda]
"""
import lzma
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT = "/root/reference/data/small_panel.preprocessed.rda"

NILSXP, SYMSXP, LISTSXP, CLOSXP, ENVSXP, PROMSXP, LANGSXP, SPECIALSXP, BUILTINSXP, CHARSXP = range(10)
LGLSXP, INTSXP, REALSXP, CPLXSXP, STRSXP, DOTSXP, VECSXP, EXPRSXP, BCODESXP = 10, 13, 14, 15, 16, 17, 19, 20, 21
EXTPTRSXP, WEAKREFSXP, RAWSXP, S4SXP = 22, 23, 24, 25
REFSXP, NILVALUE, GLOBALENV, UNBOUND, MISSINGARG, BASENAMESPACE, NAMESPACESXP, PACKAGESXP, PERSISTSXP = (
    255, 254, 253, 252, 251, 250, 249, 248, 247)
EMPTYENV, BASEENV, ATTRLANGSXP, ATTRLISTSXP, BCREPDEF, BCREPREF, ALTREP_SXP = 242, 241, 240, 239, 244, 243, 238
NA_INT = -2147483648


class RObj:
    """Anything that carries attributes: vectors, S4 objects, closures, environments."""
    __slots__ = ("kind", "value", "attrs")

    def __init__(self, kind, value=None, attrs=None):
        self.kind, self.value, self.attrs = kind, value, attrs or {}

    def attr(self, name, default=None):
        return self.attrs.get(name, default)

    def __repr__(self):
        return f"RObj({self.kind}, attrs={list(self.attrs)})"


class Reader:
    def __init__(self, buf):
        self.b, self.o, self.refs = buf, 0, []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def raw(self, n):
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def length(self):
        n = self.i32()
        if n == -1:
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + lo
        return n

    def string_vec(self):
        assert self.i32() == 0
        return [self.item() for _ in range(self.i32())]

    @staticmethod
    def pairlist_to_attrs(pl):
        return {t: v for t, v in pl} if isinstance(pl, list) else {}

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
        if t == NILVALUE:
            return None
        if t in (GLOBALENV, BASEENV, EMPTYENV, BASENAMESPACE):
            return RObj("env", {})
        if t in (UNBOUND, MISSINGARG):
            return RObj("missing")
        if t == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t in (NAMESPACESXP, PACKAGESXP, PERSISTSXP):
            o = RObj("namespace", self.string_vec())
            self.refs.append(o)
            return o
        if t == SYMSXP:
            name = self.item()
            self.refs.append(name)
            return name
        if t == ENVSXP:
            env = RObj("env", {})
            self.i32()  # locked
            self.refs.append(env)
            self.item()  # enclosure
            frame, hashtab, attrib = self.item(), self.item(), self.item()
            if isinstance(frame, list):
                for tag, v in frame:
                    env.value[tag] = v
            if isinstance(hashtab, RObj) and hashtab.kind == "list":
                for chain in hashtab.value:
                    if isinstance(chain, list):
                        for tag, v in chain:
                            env.value[tag] = v
            env.attrs = self.pairlist_to_attrs(attrib)
            return env
        if t in (LISTSXP, LANGSXP, CLOSXP, PROMSXP, DOTSXP, ATTRLANGSXP, ATTRLISTSXP):
            # iterative over the CDR chain of pairlists
            out = []
            first_kind = t
            while True:
                attrib = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                if first_kind in (CLOSXP, PROMSXP):
                    cdr = self.item()
                    o = RObj("closure" if first_kind == CLOSXP else "promise", (tag, car, cdr),
                             self.pairlist_to_attrs(attrib))
                    return o
                nflags = self.i32()
                nt = nflags & 0xFF
                if nt in (LISTSXP, LANGSXP, DOTSXP):
                    has_attr, has_tag = bool(nflags & 0x200), bool(nflags & 0x400)
                    continue
                # the CDR is not a plain cons cell: rewind and read it as an item (normally NILVALUE)
                self.o -= 4
                tail = self.item()
                if isinstance(tail, list):
                    out.extend(tail)
                return out
        if t in (SPECIALSXP, BUILTINSXP):
            return RObj("builtin", self.raw(self.i32()).decode())
        if t == CHARSXP:
            n = self.i32()
            return None if n == -1 else self.raw(n).decode("utf-8", "replace")
        if t == EXTPTRSXP:
            o = RObj("extptr")
            self.refs.append(o)
            self.item()
            self.item()
        elif t == WEAKREFSXP:
            o = RObj("weakref")
            self.refs.append(o)
        elif t in (LGLSXP, INTSXP):
            n = self.length()
            o = RObj("int" if t == INTSXP else "lgl", np.frombuffer(self.raw(4 * n), ">i4").astype(np.int32))
        elif t == REALSXP:
            n = self.length()
            o = RObj("real", np.frombuffer(self.raw(8 * n), ">f8").astype(np.float64))
        elif t == CPLXSXP:
            n = self.length()
            o = RObj("cplx", np.frombuffer(self.raw(16 * n), ">f8").astype(np.float64))
        elif t == STRSXP:
            n = self.length()
            o = RObj("str", [self.item() for _ in range(n)])
        elif t in (VECSXP, EXPRSXP):
            n = self.length()
            o = RObj("list", [self.item() for _ in range(n)])
        elif t == RAWSXP:
            o = RObj("raw", self.raw(self.length()))
        elif t == S4SXP:
            o = RObj("S4")
        elif t == BCODESXP:
            reps = [None] * self.i32()
            o = RObj("bytecode", self.bc1(reps))
        elif t == ALTREP_SXP:
            info, state, attr = self.item(), self.item(), self.item()
            o = self.altrep(info, state)
            o.attrs.update(self.pairlist_to_attrs(attr))
            return o
        else:
            raise ValueError(f"unsupported SEXP type {t} at offset {self.o}")
        if has_attr:
            o.attrs = self.pairlist_to_attrs(self.item())
        return o

    def bc1(self, reps):
        code = self.item()
        consts = []
        for _ in range(self.i32()):
            ct = self.i32()
            if ct == BCODESXP:
                consts.append(self.bc1(reps))
            elif ct in (LANGSXP, LISTSXP, BCREPDEF, BCREPREF, ATTRLANGSXP, ATTRLISTSXP):
                consts.append(self.bclang(ct, reps))
            else:
                consts.append(self.item())
        return (code, consts)

    def bclang(self, t, reps):
        if t == BCREPREF:
            return reps[self.i32()]
        if t in (BCREPDEF, LANGSXP, LISTSXP, ATTRLANGSXP, ATTRLISTSXP):
            pos = -1
            if t == BCREPDEF:
                pos, t = self.i32(), self.i32()
            cell = ["cons", None, None, None]
            if pos >= 0:
                reps[pos] = cell
            if t in (ATTRLANGSXP, ATTRLISTSXP):
                self.item()
            cell[1] = self.item()
            cell[2] = self.bclang(self.i32(), reps)
            cell[3] = self.bclang(self.i32(), reps)
            return cell
        return self.item()

    def altrep(self, info, state):
        cls = info[0][1] if isinstance(info, list) else str(info)
        if cls in ("compact_intseq", "compact_realseq"):
            n, start, step = state.value[:3]
            seq = start + step * np.arange(int(n))
            return RObj("int", seq.astype(np.int32)) if cls == "compact_intseq" else RObj("real", seq.astype(np.float64))
        if cls.startswith("wrap_"):
            inner = state.value[0] if isinstance(state, RObj) and state.kind == "list" else state[0][1]
            return RObj(inner.kind, inner.value, dict(inner.attrs))
        if cls == "deferred_string":
            arg = state[0][1] if isinstance(state, list) else state
            vals = arg.value
            return RObj("str", [None if (isinstance(v, (int, np.integer)) and v == NA_INT) else
                                (str(int(v)) if float(v).is_integer() else repr(float(v))) for v in vals])
        raise ValueError(f"unsupported ALTREP class {cls}")


def read_rda(path):
    raw = open(path, "rb").read()
    try:
        buf = lzma.decompress(raw)
    except lzma.LZMAError:
        import bz2
        import gzip
        try:
            buf = gzip.decompress(raw)
        except OSError:
            buf = bz2.decompress(raw)
    assert buf[:5] == b"RDX3\n" and buf[5:7] == b"X\n", buf[:8]
    r = Reader(buf)
    r.o = 7
    version, _, _ = r.i32(), r.i32(), r.i32()
    assert version == 3
    r.raw(r.i32())  # native encoding
    top = r.item()
    return {tag: v for tag, v in top}


def dgc(m):
    """S4 dgCMatrix -> (indptr, indices, data, shape, rownames, colnames)"""
    a = m.attrs
    dn = a["Dimnames"].value
    return (a["p"].value.astype(np.int32), a["i"].value.astype(np.int32), a["x"].value.astype(np.float64),
            tuple(int(v) for v in a["Dim"].value), dn[0].value if dn[0] is not None else None,
            dn[1].value if dn[1] is not None else None)


def named_list(o):
    names = o.attr("names")
    return dict(zip(names.value, o.value)) if names is not None else {}


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else DEFAULT
    top = read_rda(path)
    panel = top["small_panel.preprocessed"]
    samples = named_list(panel)
    out = {"sample_names": np.array(list(samples), dtype=object)}
    for name, p2 in samples.items():
        env = p2.value  # the R6 object is an environment
        indptr, indices, data, shape, rn, cn = dgc(env["counts"])   # cells x genes, CSC by gene (Pagoda2$counts)
        misc = named_list(env["misc"])
        varinfo = misc["varinfo"]
        vi = named_list(varinfo)
        vrows = varinfo.attr("row.names")
        reds = named_list(env["reductions"])
        pca = reds["PCA"]
        pdim = tuple(int(v) for v in pca.attr("dim").value)
        # Pagoda2$getOdGenes(n.odgenes) [3P, pagoda2 is not vendored in the reference tree]:
        #   rownames(varinfo)[order(varinfo$lp, decreasing = FALSE)[1:min(ncol(counts), n.odgenes)]]
        # R's order() is stable and puts NaN last; conos always passes n.odgenes (R/conos.R:108), so the list of
        # all genes in that order is what the mirror slices.
        lp = vi["lp"].value
        vn_all = list(varinfo.attr("row.names").value)
        order = np.argsort(np.where(np.isnan(lp), np.inf, lp), kind="stable")
        od = [vn_all[t] for t in order]
        # varinfo rows follow the gene order of `counts` in pagoda2; map by name to be safe
        vnames = list(vrows.value) if vrows is not None and vrows.kind == "str" else list(cn)
        pos = {g: k for k, g in enumerate(vnames)}
        idx = np.array([pos[g] for g in cn])
        out[f"{name}/indptr"], out[f"{name}/indices"], out[f"{name}/data"] = indptr, indices, data
        out[f"{name}/shape"] = np.array(shape, np.int64)
        out[f"{name}/cells"] = np.array(rn, dtype=object)
        out[f"{name}/genes"] = np.array(cn, dtype=object)
        out[f"{name}/varinfo_v"] = vi["v"].value[idx]
        out[f"{name}/varinfo_qv"] = vi["qv"].value[idx]
        out[f"{name}/pca"] = pca.value.reshape(pdim[1], pdim[0]).T.copy()   # column-major -> [cells][pcs]
        out[f"{name}/od_genes"] = np.array(od, dtype=object)
        print(name, "counts", shape, "nnz", len(data), "pca", pdim, "od genes", len(od))
    dst = os.path.join(HERE, "small_panel_preprocessed.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
