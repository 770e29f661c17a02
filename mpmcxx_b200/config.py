"""Hot-path options: the subset of the reference's input keywords that parameterise System::energy().

Mirrors src/SimulationControl.cpp:258-1616 (`process_command`) for those keywords only, with the reference's
defaults (src/System.h:21-22 ewald_kmax 7; :631 rd_lrc 1; :698 polar_gamma 1.0; alphas derived as 3.5/cutoff
unless set, src/System.cpp:871-874).  Used to fill the C-ABI's `mpmc_config` and the oracle's option arrays.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict

DAMPING = {"off": 0, "none": 0, "linear": 1, "exponential": 2}


@dataclass
class EnergyOptions:
    rd_lrc: int = 1
    rd_only: int = 0
    polarization: int = 0
    damp_type: int = 0
    polar_ewald: int = 0
    polar_iterative: int = 0
    polar_gs: int = 0
    polar_gs_ranked: int = 0
    polar_palmo: int = 0
    polar_sor: int = 0
    polar_esor: int = 0
    polar_zodid: int = 0
    polar_rrms: int = 0
    polar_max_iter: int = 0
    ewald_kmax: int = 7
    polar_damp: float = 0.0
    polar_gamma: float = 1.0
    polar_precision: float = 0.0
    ewald_alpha: float = 0.0          # <= 0: derive 3.5/cutoff
    polar_ewald_alpha: float = 0.0    # <= 0: derive 3.5/cutoff
    temperature: float = 0.0

    def as_dict(self):
        return asdict(self)


_ONOFF = ("rd_lrc", "rd_only", "polarization", "polar_ewald", "polar_iterative", "polar_gs", "polar_gs_ranked",
          "polar_palmo", "polar_sor", "polar_esor", "polar_zodid", "polar_rrms")
_INT = ("polar_max_iter", "ewald_kmax")
_FLOAT = ("polar_damp", "polar_gamma", "polar_precision", "ewald_alpha", "polar_ewald_alpha", "temperature")


def from_keywords(opts: dict) -> EnergyOptions:
    """Translate input-file keywords (as strings) to EnergyOptions; unknown keywords are ignored here because
    they do not touch the energy path (they belong to the MC driver / IO)."""
    o = EnergyOptions()
    for k, v in opts.items():
        v = str(v).strip()
        if k in _ONOFF:
            lv = v.lower()
            if lv not in ("on", "off"):
                raise ValueError("invalid setting for %s: %r" % (k, v))   # reference: return fail -> invalid_input
            setattr(o, k, 1 if lv == "on" else 0)
        elif k in _INT:
            setattr(o, k, int(v))
        elif k in _FLOAT:
            setattr(o, k, float(v))
        elif k == "polar_damp_type":
            if v.lower() not in DAMPING:
                raise ValueError("invalid polar_damp_type %r" % v)
            o.damp_type = DAMPING[v.lower()]
    return o
