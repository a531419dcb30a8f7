#!/bin/bash
# Regenerates profiles/sass/*.sass (cuobjdump -sass of the hot kernels in libngp_b200.so) — the instantiations the
# headline step launches: density net (feature tiles) fw<6 slots>/bw<5>, colour net fw<8>/bw<4>, hash grid F=2 fp32, marcher, compositors.
set -e
cd "$(dirname "$0")/../profiles/sass"
SO="../../instant-ngp-pp_b200/libngp_b200.so"
cuobjdump -sass $SO | grep "Function :" | sed 's/.*Function : //' > /tmp/fn_list.txt
rm -f mlp_*.sass hashgrid_*.sass density_head_*.sass density_net_*.sass get_rays.sass
dump() {  # pattern outname
  fn=$(grep -E "$1" /tmp/fn_list.txt | head -1)
  [ -z "$fn" ] && { echo "no function matches $1"; return; }
  cuobjdump -sass -fun "$fn" $SO 2>/dev/null | sed 's#/\*[0-9a-f]\{4\}\*/##; s#/\* 0x[0-9a-f]* \*/##' | grep -v "^\s*$" > "$2.sass"
}
dump 'mlp_fw_kernelILi6ENS_11StaticShapeILi1ELi2E' mlp_fw_sigma_tiles_slots6
dump 'mlp_fw_kernelILi8ENS_11StaticShapeILi2E' mlp_fw_rgb_slots8
dump 'mlp_bw_kernelILi5ENS_11StaticShapeILi1ELi2E' mlp_bw_sigma_tiles_slots5
dump 'mlp_bw_kernelILi4ENS_11StaticShapeILi2E' mlp_bw_rgb_slots4
dump 'hashgrid_fw_kernelILi2EfLb1E' hashgrid_fw_tiles_F2
dump 'hashgrid_bw_params_kernelILi2ELi2ELb1E' hashgrid_bw_params_tiles_F2
dump 'hashgrid_bw_params_f8_kernelILb0ELi0E' hashgrid_bw_params_f8
dump 'hashgrid_bw_params_f8_kernelILb0ELi1E' hashgrid_bw_params_f8_double_backward
dump 'hashgrid_bw_params_f8_kernelILb0ELi2E' hashgrid_bw_params_f8_dual
dump 'hashgrid_fw_kernelILi8EfLb0ELb1E' hashgrid_fw_F8_double_backward
dump 'density_net_fw_kernel' density_net_fw
dump 'density_net_bw_kernel' density_net_bw
dump 'density_head_fw_kernelILi1E' density_head_fw_W128
dump 'density_head_bw_kernelILi1E' density_head_bw_W128
dump 'get_rays_kernel' get_rays
dump 'march_count_kernelILb1E' march_count_kernelILb1E
dump 'composite_train_fw_kernelILi32E' composite_train_fw_kernelILi32E
dump 'composite_train_bw_kernelILi32E' composite_train_bw_kernelILi32E
grep -c "UTCHMMA" mlp_*.sass density_net_*.sass; grep -c "UBLKCP" mlp_*.sass; grep -c "LDTM\|STTM" mlp_*.sass
