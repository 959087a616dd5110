"""
Input-format helpers.

The reference stores preprocessed data as three contiguous int8 HDF5 datasets (`DataProcesser.py:157-161`) and its
loader widens them to float32 on the CPU (`BlockDatasetLoader.py:139-143`), so every batch crosses PCIe 4x larger
than it has to.  The drop-in modules accept int8 / uint8 `nodes` and `edges` directly -- copy the int8 batch to the
device and widen there -- and `read_hdf5_raw` reads the reference's files without h5py.
"""
import numpy as np
import torch


def read_hdf5_raw(path, max_n_nodes, n_node_features, n_edge_features, apd_len):
    """(nodes, edges, apds) int8 arrays from a GraphINVENT `*.h5` written by `resave_datasets_unchunked`
    (`DataProcesser.py:147-165`): HDF5 superblock v0, 2048-byte header, contiguous datasets in alphabetical order
    APDs / edges / nodes.  Raises if the file size does not match that layout."""
    N, F, E = max_n_nodes, n_node_features, n_edge_features
    raw = np.fromfile(path, np.int8)
    row = apd_len + N * N * E + N * F
    n, rem = divmod(raw.size - 2048, row)
    if n <= 0 or rem != 0:
        raise ValueError(f"{path}: size {raw.size} is not 2048 + n*{row} bytes -- not the contiguous int8 layout")
    o = 2048
    apds = raw[o:o + n * apd_len].reshape(n, apd_len); o += n * apd_len
    edges = raw[o:o + n * N * N * E].reshape(n, N, N, E); o += n * N * N * E
    nodes = raw[o:o + n * N * F].reshape(n, N, F)
    return nodes, edges, apds


def to_device_batch(nodes, edges, apds=None, device="cuda"):
    """pinned int8 host batch -> device (1 byte per element over PCIe); the modules widen to fp32 on the device"""
    out = []
    for a in (nodes, edges, apds):
        if a is None:
            continue
        t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
        out.append(t.pin_memory().to(device, non_blocking=True) if t.device.type == "cpu" else t)
    return out
