"""cnld.arrays -- array generators (mirror of /root/reference/cnld/scripts/matrix_array.py:10-164
and cnld/_arrays/matrix.py; same defaults, positions and patch numbering)."""
import numpy as np

from .abstract import Array, Element, Patch, SquareCmutMembrane


def _centered_grid(n, pitch):
    xx, yy, zz = np.meshgrid(np.linspace(0, (n[0] - 1) * pitch[0], n[0]),
                             np.linspace(0, (n[1] - 1) * pitch[1], n[1]), 0)
    return np.c_[xx.ravel(), yy.ravel(), zz.ravel()] - [(n[0] - 1) * pitch[0] / 2, (n[1] - 1) * pitch[1] / 2, 0]


def matrix_array(nelem=(2, 2), elempitch=(60e-6, 60e-6), nmem=(1, 1), mempitch=(60e-6, 60e-6),
                 length=(40e-6, 40e-6), electrode=(40e-6, 40e-6), npatch=(3, 3), thickness=(2e-6,),
                 density=(2040,), y_modulus=(110e9,), p_ratio=(0.22,), isolation=200e-9, permittivity=6.3,
                 gap=50e-9, damping_mode_a=0, damping_mode_b=4, damping_ratio_a=0.1, damping_ratio_b=0.1):
    """Square-membrane matrix array with the defaults of scripts/matrix_array.py:168-199."""
    ppx, ppy = length[0] / npatch[0], length[1] / npatch[1]
    patch_pos = _centered_grid(npatch, (ppx, ppy))
    mem_pos = _centered_grid(nmem, mempitch)
    elem_pos = _centered_grid(nelem, elempitch)
    props = dict(length_x=length[0], length_y=length[1], electrode_x=electrode[0], electrode_y=electrode[1],
                 y_modulus=tuple(y_modulus), p_ratio=tuple(p_ratio), isolation=isolation,
                 permittivity=permittivity, gap=gap, damping_mode_a=damping_mode_a,
                 damping_mode_b=damping_mode_b, damping_ratio_a=damping_ratio_a,
                 damping_ratio_b=damping_ratio_b, thickness=tuple(thickness), density=tuple(density))
    elements = []
    pc = mc = 0
    for ec, epos in enumerate(elem_pos):
        membranes = []
        for mpos in mem_pos:
            patches = []
            for ppos in patch_pos:
                patches.append(Patch(id=pc, position=(epos + mpos + ppos).tolist(), length_x=ppx, length_y=ppy,
                                     area=ppx * ppy))
                pc += 1
            membranes.append(SquareCmutMembrane(id=mc, position=(epos + mpos).tolist(), patches=patches, **props))
            mc += 1
        elements.append(Element(id=ec, position=epos.tolist(), kind='both', membranes=membranes))
    return Array(id=0, position=[0, 0, 0], elements=elements)
