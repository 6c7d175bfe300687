"""``$Frag`` input parser -- the ``solvshift.md.MDInput`` surface (solvshift/md.py:32-196).

    $Frag
      NMA        nma
      reorder    1 3 5 2 7 10 11 12 4 6 8 9
      supl       1-5
      atoms      1-12                           1
    $endFrag

``MDInput(text_or_file).get()`` returns ``args = [ind, nmol, bsm, supl, reord]`` for ``EFP.set`` and
``idx``, the flat atom slice into an MD frame.  Lists are converted to 0-based on reading.
"""
import os
import re

import numpy

from .slvpar import Frag, _text_to_list

__all__ = ['MDInput']


class MDInput(object):
    def __init__(self, file_name):
        self._file_name = file_name
        self._ind = []; self._nmol = []; self._bsm = []; self._supl = []; self._reord = []
        self._frame_slice = []
        self._frag_idx = 0
        self._type_of = {}
        self._input_text = None
        self(file_name)

    def get(self):
        """args, idx = MDInputInstance.get()"""
        return self.get_args(), self._frame_slice

    def get_idx(self):
        return self._frame_slice

    def get_args(self):
        return [self._ind, self._nmol, self._bsm, self._supl, self._reord]

    def get_efp(self, frame):
        """parse EFP coordinates from MD frame"""
        return frame[self._frame_slice]

    def __call__(self, file_name):
        if os.path.exists(file_name):
            with open(file_name) as fh:
                text = fh.read()
        elif '\n' in file_name:
            text = file_name
        else:
            raise IOError(" The input file does not exists.")
        self._input_text = text
        for section in re.split(r'\$Frag', text)[1:]:
            self._read(section)

    def _read(self, section):
        section = section.split('$endFrag')[0]
        lines = [ln.strip() for ln in section.split('\n')[1:]]
        lines = [ln for ln in lines if ln]
        if not lines or lines[0].startswith('!'):
            return
        frag_name = lines[0].split()[1]
        reord = None; supl = None; n_atoms_per_mol = None; spans = []
        for line in lines[1:]:
            if line.startswith('!'):
                continue
            if line.startswith('reorder'):
                reord = numpy.array([int(x) for x in line.split()[1:]], int) - 1
            elif line.startswith('supl'):
                ll = line.split('supl')[-1]
                supl = _text_to_list(ll, delimiter=',' if ',' in ll else None) - 1
            elif line.startswith('atoms'):
                atoms, n_frags = line.split()[1:]
                atoms = _text_to_list(atoms, delimiter=',' if ',' in atoms else None)
                n_frags = int(n_frags)
                n_atoms = atoms[-1] - atoms[0] + 1
                assert n_atoms % n_frags == 0, \
                    'MDInputError: Invalid atomic indices or fragment numbers in fragment %i\n -----> %s' % (self._frag_idx + 1, line)
                n_atoms_per_mol = n_atoms // n_frags
                spans.append((n_frags, n_atoms_per_mol))
                self._frame_slice += list(numpy.array(atoms) - 1)
        if reord is None:
            reord = numpy.arange(n_atoms_per_mol, dtype=int)
        # A protein input (biomol.py:249-328) repeats the same (fragment, reorder, supl) block for every residue of a kind:
        # such blocks share ONE fragment type (the reference appends a freshly parsed Frag per block; `ind` then simply
        # points at the shared entry, which changes nothing for EFP.set)
        key = (frag_name, tuple(int(x) for x in reord), None if supl is None else tuple(int(x) for x in supl))
        t = self._type_of.get(key)
        if t is None:
            t = len(self._bsm); self._type_of[key] = t
            self._bsm.append(Frag(frag_name)); self._supl.append(supl); self._reord.append(reord)
        for n_frags, per in spans:
            for i in range(n_frags):
                self._ind.append(t); self._nmol.append(per)
        self._frag_idx += 1

    def __repr__(self):
        N = 50
        return '-' * N + '\nSLV-MD input file: ' + str(self._file_name)[:60] + '\n\n' + str(self._input_text) + '-' * N + '\n'
