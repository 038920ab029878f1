"""numpy restatement of i3d_degree_plan (include/i3d.h), shared by the CPU host-logic test and the GPU parity case."""
import numpy as np


def degree_plan_ref(rowptr, NB, CT):
    """numpy restatement of i3d_degree_plan (include/i3d.h): nodes grouped by in-degree into whole 128-row tiles."""
    deg = np.diff(np.asarray(rowptr, dtype=np.int64))
    N = len(deg)
    over = int((deg >= NB).any())
    d = np.minimum(deg, NB - 1)
    tiles = (N + 127) // 128
    T, CH = tiles + NB, (tiles + CT - 1) // CT + NB
    perm = -np.ones(T * 128, dtype=np.int32)
    tile_bucket = -np.ones(T, dtype=np.int32)
    chunk = np.zeros(3 * CH, dtype=np.int32)
    row0 = ch0 = 0
    for b in range(NB):
        nodes = np.nonzero(d == b)[0]
        perm[row0:row0 + len(nodes)] = nodes
        tb = (len(nodes) + 127) // 128
        tile_bucket[row0 // 128:row0 // 128 + tb] = b
        for j in range((tb + CT - 1) // CT):
            r0 = row0 + j * CT * 128
            chunk[3 * ch0:3 * ch0 + 3] = (r0, min(CT * 128, row0 + tb * 128 - r0), b)
            ch0 += 1
        row0 += tb * 128
    return perm, tile_bucket, chunk, over
