"""Shared parity cases: the BASELINE.json configs plus coverage boards.

Narrow boards are here because uniform random actions on 10x20 clear ~1 line
per 3e5 steps: widths 4..6 exercise singles..tetrises.
"""
import numpy as np

SHAPE_NAMES = ["T", "J", "L", "Z", "S", "I", "O"]

CASES = {
    # BASELINE.json configs
    "C1_default_ram": dict(),
    "C2_ram_step_adv": dict(reward_step=True, advanced_clears=True),
    "C3_phi_holesinc_ld3_sr": dict(penalise_height_increase=True, penalise_holes_increase=True,
                                   lock_delay=3, step_reset=True),
    "C4_gray_ext_high": dict(obs_type="grayscale", extend_dims=True, high_scoring=True),
    "C5a_rgb": dict(obs_type="rgb"),
    "C5b_ram_40x20": dict(width=20, height=40),
    # coverage: line clears, every flag, odd geometry
    "narrow4_adv_step": dict(width=4, height=20, reward_step=True, advanced_clears=True),
    "narrow4_default": dict(width=4, height=20),
    "narrow4_high_ph_pholes": dict(width=4, height=20, high_scoring=True, penalise_height=True,
                                   penalise_holes=True),
    "narrow5x12_ld3_sr_inc": dict(width=5, height=12, lock_delay=3, step_reset=True,
                                  penalise_height_increase=True, penalise_holes_increase=True),
    "w6h10_ld2": dict(width=6, height=10, lock_delay=2, penalise_holes_increase=True),
    "w7h9_odd_gray": dict(width=7, height=9, obs_type="grayscale"),
    "w16h16_rgb": dict(width=16, height=16, obs_type="rgb", reward_step=True),
    "w4h41_gray": dict(width=4, height=41, obs_type="grayscale", extend_dims=True),
    "w20h40_gray": dict(width=20, height=40, obs_type="grayscale"),
    "w17h31_ram_ext": dict(width=17, height=31, extend_dims=True, penalise_height=True),
    "w32h33_ram": dict(width=32, height=33, penalise_holes=True),
    "w3h8_all_flags": dict(width=3, height=8, reward_step=True, penalise_height=True,
                           penalise_height_increase=True, advanced_clears=True, high_scoring=True,
                           penalise_holes=True, penalise_holes_increase=True),
    "w4h50_rgb_degenerate": dict(width=4, height=50, obs_type="rgb"),
    "w10h20_ldneg": dict(lock_delay=-5, reward_step=True),
    "w4h20_ld1_sr": dict(width=4, height=20, lock_delay=1, step_reset=True, advanced_clears=True),
}


def actions_for(seed, T):
    """Frozen legacy stream."""
    return np.random.RandomState(seed).randint(0, 7, T).astype(np.uint8)
