#!/bin/bash
rm -f /tmp/ab2_ref.npz
L=profiles/ab/libcfear_
bash profiles/gpu_ab2.sh r03l ${L}base.so ${L}k3s.so ${L}base.so ${L}k3s.so
