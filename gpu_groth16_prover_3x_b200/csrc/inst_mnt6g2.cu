// Instantiation of the per-group host routines and kernels for Mnt6G2 (see group_ops.cuh).
#include "group_ops.cuh"

extern const GroupOps b200msm_ops_mnt6g2;
const GroupOps b200msm_ops_mnt6g2 = make_group_ops<mnt753::Mnt6G2>();
