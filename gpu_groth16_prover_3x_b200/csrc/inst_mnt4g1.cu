// Instantiation of the per-group host routines and kernels for Mnt4G1 (see group_ops.cuh).
#include "group_ops.cuh"

extern const GroupOps b200msm_ops_mnt4g1;
const GroupOps b200msm_ops_mnt4g1 = make_group_ops<mnt753::Mnt4G1>();
