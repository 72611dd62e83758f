"""`python -m deepqnetwork_b200.formats <save_dir>/<environment>`: list what a save-dir prefix holds, as this build reads it.

Meant for the first contact with a save-dir written by the reference itself (TensorFlow + h5py): the format code here has
never met such files, so look before resuming from one.  Every table block and every tensor is CRC-checked;
unsupported structures are reported by name.  Host-side only, no GPU needed."""
import os
import sys

import numpy as np

from . import hdf5, tf_bundle


def describe(prefix, out=sys.stdout):
    ok = True
    if os.path.exists(prefix + '.index'):
        r = tf_bundle.BundleReader(prefix)
        out.write('checkpoint %s.index: %d variables, header %s\n' % (prefix, len(r.names()), r.header))
        total = 0
        for name in r.names():
            e = r.entry(name)
            try:
                t = r.tensor(name)
                note = 'crc ok' + ('  = %s' % t if t.shape == () else '')
            except Exception as exc:  # report and go on: the point is to see everything that is wrong
                ok = False
                note = 'UNREADABLE: %s' % exc
            total += e['size']
            out.write('  %-64s %-8s %-18s %s\n' % (name, tf_bundle.DT_INV.get(e['dtype'], 'dtype#%d' % e['dtype']), e['shape'], note))
        out.write('  %d bytes of tensor data\n' % total)
    else:
        out.write('no TensorFlow checkpoint at %s.index\n' % prefix)
    path = prefix + '_replay_memory.hdf5'
    if os.path.exists(path):
        with hdf5.File(path) as f:
            out.write('replay file %s\n' % path)
            for k, v in f.attrs.items():
                out.write('  attr %-16s = %r\n' % (k, v))
            for k, d in sorted(f.datasets.items()):
                where = 'compact' if d.compact is not None else ('not allocated (reads as zeros)' if not d.allocated else 'at %d' % d.address)
                out.write('  dataset %-12s %-8s %-22s %s\n' % (k, np.dtype(d.dtype).name, d.shape, where))
    else:
        out.write('no replay file at %s\n' % path)
    if os.path.exists(prefix + '.dqnb200'):
        out.write('native container %s.dqnb200 present (used only when there is no .index)\n' % prefix)
    return ok


if __name__ == '__main__':
    if len(sys.argv) != 2:
        sys.exit('usage: python -m deepqnetwork_b200.formats <save_dir>/<environment dotted path>')
    sys.exit(0 if describe(sys.argv[1]) else 1)
