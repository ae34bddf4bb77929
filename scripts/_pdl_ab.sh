for i in 1 2 3; do
for p in 1 0; do
MG_PDL=$p python - <<'PY'
import os, sys, json, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'scripts')
import bench_sections as S
torch.cuda.set_device(0)
pk = {'hbm_gbs': 6549.8, 'bf16_tflops': 1670.5}
t = S.training_section(0, 1, torch.device('cuda', 0), pk)
print('MG_PDL=%s graph %.4f ms eager %.4f ms without-allreduce %.4f' % (os.environ['MG_PDL'], t['ms_per_step'], t['eager_ms_per_step'], t['gradient_allreduce']['step_without_it_ms']))
PY
done; done
