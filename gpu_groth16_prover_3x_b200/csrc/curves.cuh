// The four groups of the MNT4753 / MNT6753 Groth16 prover (reference: multiexp/curves.cu:421-425,
// libff mnt4753_init.cpp:105-128, mnt6753_init.cpp:109-142).
#pragma once
#include "ec.cuh"

namespace mnt753 {

// scalars of a curve live in the *other* curve's base field (curves.cu `group_type`)
struct Mnt4G1 { typedef FieldCfg<ModA, 1, 1, 0, 2, 1> F;    typedef ModB Fr; static constexpr int CURVE = 0, GROUP = 1; };
struct Mnt4G2 { typedef FieldCfg<ModA, 2, 13, 1, 26, 1> F;  typedef ModB Fr; static constexpr int CURVE = 0, GROUP = 2; };
struct Mnt6G1 { typedef FieldCfg<ModB, 1, 1, 0, 11, 1> F;   typedef ModA Fr; static constexpr int CURVE = 1, GROUP = 1; };
struct Mnt6G2 { typedef FieldCfg<ModB, 3, 11, 2, 11, 121> F; typedef ModA Fr; static constexpr int CURVE = 1, GROUP = 2; };

}  // namespace mnt753
