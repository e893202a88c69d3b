// Instantiation of the per-group host routines and kernels for Mnt4G2 (see group_ops.cuh).
#include "group_ops.cuh"

extern const GroupOps b200msm_ops_mnt4g2;
const GroupOps b200msm_ops_mnt4g2 = make_group_ops<mnt753::Mnt4G2>();
