"""Loading and saving of AstroLink instances: same functions, signatures and file format as the reference's
astrolink/io.py (saveAstroLinkObject :12-24, loadAstroLinkObject :26-46) -- one pickled object under the key
'clusterer' of an .npz file.  Results that still live on the GPU are copied to numpy arrays when the object is
pickled (AstroLink.__getstate__), so a saved object holds no CUDA handles and loads on a machine without a GPU;
every attribute the reference's plotting functions read (visualize.py: logRho, ordering, clusters, ids, groups,
prominences, groups_sigs, pFit, S, n_samples) is a plain numpy array / scalar after loading.
"""
import numpy as np


def saveAstroLinkObject(clusterer, fileName):
    """Save an AstroLink instance (reference: io.py:12-24)."""
    np.savez(fileName, clusterer=clusterer)


def loadAstroLinkObject(fileName):
    """Load a saved AstroLink instance (reference: io.py:26-46)."""
    if not fileName.endswith('.npz'):
        fileName += '.npz'
    return np.load(fileName, allow_pickle=True)['clusterer'].item()
