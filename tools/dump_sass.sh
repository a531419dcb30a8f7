#!/bin/bash
# Regenerates profiles/sass/*.sass (cuobjdump -sass of the hot kernels in libngp_b200.so).
set -e
cd "$(dirname "$0")/../profiles/sass"
SO="../../instant-ngp-pp_b200/libngp_b200.so"
cuobjdump -sass $SO | grep "Function :" | sed 's/.*Function : //' > /tmp/fn_list.txt
for pat in mlp_fw_kernel mlp_bw_kernel hashgrid_fw_kernelILi2EfE hashgrid_bw_params_kernelILi2E march_count_kernelILb1E composite_train_fw_kernelILi16E composite_train_bw_kernelILi16E; do
  fn=$(grep "$pat" /tmp/fn_list.txt | head -1)
  cuobjdump -sass -fun "$fn" $SO | sed 's#/\*[0-9a-f]\{4\}\*/##; s#/\* 0x[0-9a-f]* \*/##' | grep -v "^\s*$" > "$pat.sass"
done
