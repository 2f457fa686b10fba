"""Plain-data description of a stream setup (what StreamDecoder.LoadBooks produces, StreamDecoder.cs:226-289)
and its .npz serialisation.  A description is a dict:
  channels, sample_rate, block_size (2), books [{dims, entries, map_type, table}], floors [{type, n_posts,
  multiplier, range, x_list, l_neigh, h_neigh, sort_idx}], residues [{type, begin, end, partition_size,
  classifications, max_stages, cascade, books[class][stage]}], mappings [{n_coupling, n_submaps, floor,
  residue, magnitude, angle}], modes [{block_flag, mapping}]
"""
from __future__ import annotations

import numpy as np

from . import capi


def to_setup(desc: dict) -> capi.Setup:
    return capi.Setup(desc["channels"], desc["sample_rate"], desc["block_size"], desc["books"], desc["floors"],
                      desc["residues"], desc["mappings"], desc["modes"])


def save(path: str, desc: dict, **extra):
    z = dict(extra)
    z["hdr"] = np.array([desc["channels"], desc["sample_rate"], desc["block_size"][0], desc["block_size"][1], len(desc["books"]),
                         len(desc["floors"]), len(desc["residues"]), len(desc["mappings"]), len(desc["modes"])], np.int64)
    z["book_info"] = np.array([[b["dims"], b["entries"], b["map_type"], 0 if b.get("table") is None else len(b["table"])] for b in desc["books"]], np.int64)
    tabs = [np.asarray(b["table"], np.float32) for b in desc["books"] if b.get("table") is not None and len(b["table"])]
    z["book_tables"] = np.concatenate(tabs) if tabs else np.zeros(0, np.float32)
    fl = np.zeros((len(desc["floors"]), 4 + 4 * 64), np.int32)
    for i, f in enumerate(desc["floors"]):
        n = int(f.get("n_posts", 0))
        fl[i, :4] = [f["type"], n, f.get("multiplier", 0), f.get("range", 0)]
        for j, key in enumerate(("x_list", "l_neigh", "h_neigh", "sort_idx")):
            if n:
                fl[i, 4 + 64 * j: 4 + 64 * j + n] = f[key][:n]
    z["floors"] = fl
    rs = np.full((len(desc["residues"]), 6 + 64 + 512), -1, np.int32)
    for i, r in enumerate(desc["residues"]):
        nc = int(r["classifications"])
        rs[i, :6] = [r["type"], r["begin"], r["end"], r["partition_size"], nc, r["max_stages"]]
        rs[i, 6:6 + 64] = 0
        rs[i, 6:6 + nc] = r["cascade"][:nc]
        bk = np.full((64, 8), -1, np.int32)
        for c in range(nc):
            row = list(r["books"][c])[:8]
            bk[c, :len(row)] = row
        rs[i, 70:] = bk.reshape(-1)
    z["residues"] = rs
    mp = np.zeros((len(desc["mappings"]), 4 + 64), np.int32)
    for i, m in enumerate(desc["mappings"]):
        n = int(m["n_coupling"])
        mp[i, :4] = [n, m["n_submaps"], m["floor"], m["residue"]]
        mp[i, 4:4 + n] = m["magnitude"][:n]
        mp[i, 36:36 + n] = m["angle"][:n]
    z["mappings"] = mp
    z["modes"] = np.array([[m["block_flag"], m["mapping"]] for m in desc["modes"]], np.int32)
    np.savez_compressed(path, **z)


def load(path: str):
    """Returns (desc, extra arrays dict)."""
    z = dict(np.load(path))
    hdr = z.pop("hdr")
    info, tabs = z.pop("book_info"), z.pop("book_tables")
    books, off = [], 0
    for d, e, mt, tl in info:
        books.append(dict(dims=int(d), entries=int(e), map_type=int(mt), table=tabs[off:off + int(tl)].copy() if tl else None))
        off += int(tl)
    floors = []
    for row in z.pop("floors"):
        n = int(row[1])
        floors.append(dict(type=int(row[0]), n_posts=n, multiplier=int(row[2]), range=int(row[3]), x_list=row[4:4 + n].copy(),
                           l_neigh=row[68:68 + n].copy(), h_neigh=row[132:132 + n].copy(), sort_idx=row[196:196 + n].copy()))
    residues = []
    for row in z.pop("residues"):
        nc = int(row[4])
        residues.append(dict(type=int(row[0]), begin=int(row[1]), end=int(row[2]), partition_size=int(row[3]), classifications=nc,
                             max_stages=int(row[5]), cascade=row[6:6 + nc].copy(), books=row[70:].reshape(64, 8)[:nc].copy()))
    mappings = []
    for row in z.pop("mappings"):
        n = int(row[0])
        mappings.append(dict(n_coupling=n, n_submaps=int(row[1]), floor=int(row[2]), residue=int(row[3]), magnitude=row[4:4 + n].copy(),
                             angle=row[36:36 + n].copy()))
    modes = [dict(block_flag=int(a), mapping=int(b)) for a, b in z.pop("modes")]
    desc = dict(channels=int(hdr[0]), sample_rate=int(hdr[1]), block_size=(int(hdr[2]), int(hdr[3])), books=books, floors=floors,
                residues=residues, mappings=mappings, modes=modes)
    return desc, z


def desc_from_setup(view) -> dict:
    """The description dict of an nvb_setup produced by the host half (hostlib.HostStream.setup()): what the workload
    generators need (structure only; codebook tables are left out)."""
    S = view.struct
    books = [dict(dims=int(S.books[i].dims), entries=int(S.books[i].entries), map_type=int(S.books[i].map_type), table=None) for i in range(S.n_books)]
    floors = []
    for i in range(S.n_floors):
        f = S.floors[i]
        n = int(f.f1.n_posts)
        floors.append(dict(type=int(f.type), n_posts=n, multiplier=int(f.f1.multiplier), range=int(f.f1.range),
                           x_list=np.array(f.f1.x_list[:n], np.int32), l_neigh=np.array(f.f1.l_neigh[:n], np.int32),
                           h_neigh=np.array(f.f1.h_neigh[:n], np.int32), sort_idx=np.array(f.f1.sort_idx[:n], np.int32),
                           order=int(f.f0.order), rate=int(f.f0.rate), bark_map_size=int(f.f0.bark_map_size), amp_bits=int(f.f0.amp_bits),
                           amp_ofs=int(f.f0.amp_ofs)))
    residues = []
    for i in range(S.n_residues):
        r = S.residues[i]
        nc = int(r.classifications)
        residues.append(dict(type=int(r.type), begin=int(r.begin), end=int(r.end), partition_size=int(r.partition_size), classifications=nc,
                             max_stages=int(r.max_stages), cascade=np.array(r.cascade[:nc], np.int32),
                             books=np.array([[int(r.books[c][k]) for k in range(8)] for c in range(nc)], np.int32)))
    mappings = []
    for i in range(S.n_mappings):
        m = S.mappings[i]
        n = int(m.n_coupling)
        mappings.append(dict(n_coupling=n, n_submaps=int(m.n_submaps), floor=int(m.floor), residue=int(m.residue),
                             magnitude=np.array(m.magnitude[:n], np.int32), angle=np.array(m.angle[:n], np.int32)))
    modes = [dict(block_flag=int(S.modes[i].block_flag), mapping=int(S.modes[i].mapping)) for i in range(S.n_modes)]
    return dict(channels=int(S.channels), sample_rate=int(S.sample_rate), block_size=(int(S.block_size[0]), int(S.block_size[1])),
                books=books, floors=floors, residues=residues, mappings=mappings, modes=modes)
