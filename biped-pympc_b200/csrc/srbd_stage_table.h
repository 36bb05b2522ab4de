// Host-side construction of the stage gather table used by PdipmWarp::load_stage_raw.
#pragma once
#include "srbd_pdipm_core.h"

namespace srbd {

inline void build_stage_slots(StageSlot* t) {
  for (int i = 0; i < SB_SIZE; ++i) { t[i].kind = ST_ZERO; t[i].rel = 0; }
  for (int r = 0; r < 12; ++r)
    for (int c = 0; c < 12; ++c) {
      const int p = a_ucol_pos(r, c);
      if (p >= 0) { t[SB_B + 14 * r + c].kind = ST_AU; t[SB_B + 14 * r + c].rel = (short)(a_ucol_off(c) + p); }
    }
  for (int c = 0; c < 12; ++c) {
    t[SB_E + c] = {ST_E, (short)c};
    const short dr = (short)(a_xcol_off(c) + a_xcol_pos(c, c));
    t[SB_AP + c] = {ST_AXPREV, dr};
    t[SB_AN + c] = {ST_AXCUR, dr};
    t[SB_HX + c] = {ST_Q, (short)c};
    t[SB_PHX + c] = {ST_QINV, (short)c};
    t[SB_HU + c] = {ST_Q, (short)(120 + c)};
    t[SB_PHXP + c] = {ST_QINVPREV, (short)c};
    t[SB_EP + c] = {ST_EPREV, (short)c};
  }
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      const short rr = (short)(a_xcol_off(6 + j) + 1 + i);   // row i of x-column 6+j
      t[SB_AP + 12 + 3 * i + j] = {ST_AXPREV, rr};
      t[SB_AN + 12 + 3 * i + j] = {ST_AXCUR, rr};
    }
    const short pv = (short)(a_xcol_off(9 + i) + 1);           // row 3+i of x-column 9+i
    t[SB_AP + 21 + i] = {ST_AXPREV, pv};
    t[SB_AN + 21 + i] = {ST_AXCUR, pv};
  }
  for (int i = 0; i < 2; ++i) t[SB_MX + i] = {ST_AU, (short)(a_ucol_off(6 + 3 * i) + A_UCOL_MX_POS)};
  for (int i = 0; i < 28; ++i) t[SB_G + i] = {ST_G, (short)i};
}

}  // namespace srbd
