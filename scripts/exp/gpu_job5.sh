./scripts/exp/red_probe
python - <<'PY'
import sys,time,torch
sys.path.insert(0,'.')
import bench, multimodal_registration_b200 as mrb
svf,img=bench.synth_inputs(32,'cpu',0)
model=mrb.voxelmorph.networks.VxmDense(bench.FULL,int_steps=7,svf_resolution=2,int_resolution=2)
a,b=img.numpy(),svf.numpy()
for i in range(3):
    t=time.perf_counter(); model.predict_deform([a,b],copy=True); print('pageable ms',1e3*(time.perf_counter()-t))
PY
