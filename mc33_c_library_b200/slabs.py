"""z-slab partition of a grid over ranks and the global index bases (host logic).

SURVEY.md section 8e: cells are independent given their corners, so the volume is
cut into contiguous z-slabs of cell layers, one per GPU.  Every grid edge / on-iso
point belongs to the slab that owns its lower sample slice (the top slab also
owns slice nz), which is the same rule that removes duplicates inside one GPU.
The only exchange is an all-gather of three integers per rank.
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class Slab:
    rank: int
    cell_z0: int      # owned cell layers [cell_z0, cell_z1)
    cell_z1: int
    z_lo: int         # sample slices the rank must hold: [z_lo, z_hi)
    z_hi: int
    is_last: bool

    @property
    def n_slices(self):
        return self.z_hi - self.z_lo


def partition(nz, world):
    """Split nz cell layers into `world` contiguous slabs (first slabs get the
    remainder).  Ranks beyond nz get no slab (None)."""
    world_eff = min(world, nz)
    base, rem = divmod(nz, world_eff)
    out, z = [], 0
    for r in range(world):
        if r >= world_eff:
            out.append(None)
            continue
        n = base + (1 if r < rem else 0)
        z0, z1 = z, z + n
        z += n
        last = z1 == nz
        # halo: one slice below (normals / on-iso neighbours of the first owned
        # slice), two above (the slice shared with the next slab is numbered by
        # that slab, and its on-iso points look one slice further)
        out.append(Slab(r, z0, z1, max(z0, 1) - 1, min(z1 + 2, nz + 1), last))
    return out


def bases(counts):
    """counts: per rank (nV, nT) in rank order (None for idle ranks).
    -> per rank (vbase, vbase_next): global id of the rank's first vertex and of
    the next rank's first vertex."""
    out, v = [], 0
    for c in counts:
        if c is None:
            out.append((v, v))
            continue
        nV, _ = c
        out.append((v, v + nV))
        v += nV
    return out
