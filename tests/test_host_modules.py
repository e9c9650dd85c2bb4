"""Host-side modules around the path (image, raw_utils, pose algebra, dataset loaders) against fixtures
produced by EXECUTING the reference's own code (tests/golden/make_golden_host.py -> host.npz) and
against the reference tests' own properties (tests/image_test.py, tests/datasets_test.py).  CPU only."""
import ast
import json
import os
import struct

import numpy as np
import pytest

from util import golden

G = golden('host')


def test_image_functions_match_reference_run():
  from multinerf_b200 import image
  np.testing.assert_allclose(image.color_correct(G['cc_img'], G['cc_ref']), G['cc_out'], atol=2e-6)   # the reference fits in float32
  np.testing.assert_allclose(image.srgb_to_linear(G['srgb_x']), G['srgb_to_linear'], rtol=1e-12, atol=0)
  np.testing.assert_allclose(image.linear_to_srgb(G['srgb_x']), G['linear_to_srgb'], rtol=1e-12, atol=0)
  np.testing.assert_array_equal(image.downsample(G['ds_in'], 4), G['ds_out'])
  np.testing.assert_allclose(image.mse_to_psnr(G['psnr_in']), G['psnr_out'], rtol=1e-6)
  with pytest.raises(ValueError):
    image.downsample(G['ds_in'], 5)
  with pytest.raises(ValueError):
    image.color_correct(G['cc_img'], G['cc_ref'][..., :2])


def test_image_reference_test_properties():
  """tests/image_test.py: colour correction undoes a CCM + quadratic warp + shift; conversions round-trip."""
  from multinerf_b200 import image
  rng = np.random.default_rng(0)
  for _ in range(3):
    im0 = rng.uniform(0.1, 0.9, (64, 64, 3))
    ccm = np.eye(3) + rng.normal(size=(3, 3)) * rng.normal() / 10
    im1 = np.clip((im0.reshape(-1, 3) @ ccm).reshape(im0.shape) + rng.normal() / 10 * im0 ** 2 + rng.normal() / 10,
                  0, 1)
    np.testing.assert_allclose(image.color_correct(im0, im1), im1, atol=1e-5, rtol=1e-5)
  for psnr in [10., 20., 30.]:
    np.testing.assert_allclose(image.mse_to_psnr(image.psnr_to_mse(psnr)), psnr, atol=1e-5, rtol=1e-5)
  for s in [-0.9, 0, 0.9]:
    np.testing.assert_allclose(image.dssim_to_ssim(image.ssim_to_dssim(s)), s, atol=1e-5)
  x = np.linspace(0, 1, 1001)
  np.testing.assert_allclose(image.linear_to_srgb(image.srgb_to_linear(x)), x, atol=1e-5)


def test_ssim_against_independent_implementation():
  """dm_pix.ssim restated (Gaussian 11 / 1.5, valid windows): checked against a scipy.ndimage evaluation."""
  from scipy import ndimage
  from multinerf_b200 import image
  rng = np.random.default_rng(1)
  a = rng.uniform(size=(40, 48, 3))
  b = np.clip(a + rng.normal(size=a.shape) * 0.1, 0, 1)
  r = np.arange(11) - 5.0
  w = np.exp(-0.5 * (r / 1.5) ** 2)
  w /= w.sum()

  def filt(x):
    y = ndimage.correlate1d(ndimage.correlate1d(x, w, axis=0, mode='constant'), w, axis=1, mode='constant')
    return y[5:-5, 5:-5]
  mu_a, mu_b = filt(a), filt(b)
  saa, sbb, sab = filt(a * a) - mu_a ** 2, filt(b * b) - mu_b ** 2, filt(a * b) - mu_a * mu_b
  c1, c2 = 0.01 ** 2, 0.03 ** 2
  ref = np.mean((2 * mu_a * mu_b + c1) * (2 * sab + c2) / ((mu_a ** 2 + mu_b ** 2 + c1) * (saa + sbb + c2)))
  assert abs(image.ssim(a, b) - ref) < 1e-9
  assert image.ssim(a, a) == pytest.approx(1.0, abs=1e-12)
  m = image.MetricHarness()(a, b, name_fn=lambda s: 'x_' + s)
  assert set(m) == {'x_psnr', 'x_ssim'} and m['x_psnr'] == pytest.approx(
      float(image.mse_to_psnr(((a - b) ** 2).mean())))


def test_raw_utils_match_reference_run():
  from multinerf_b200 import raw_utils
  np.testing.assert_array_equal(raw_utils.bilinear_demosaic(G['bayer']), G['demosaic'])     # bit-exact
  np.testing.assert_array_equal(raw_utils.pixels_to_bayer_mask(G['mask_px'], G['mask_py']), G['bayer_mask'])
  np.testing.assert_allclose(raw_utils.postprocess_raw(G['pp_raw'], G['pp_cam2rgb']), G['pp_auto'], rtol=1e-12)
  np.testing.assert_allclose(raw_utils.postprocess_raw(G['pp_raw'], G['pp_cam2rgb'], 0.6), G['pp_fixed'], rtol=1e-12)
  exifs = ast.literal_eval(str(G['exif_json'][0]))
  meta = raw_utils.process_exif(exifs)
  np.testing.assert_allclose(meta['cam2rgb'], G['exif_cam2rgb'], rtol=1e-12)
  np.testing.assert_allclose(meta['ShutterSpeed'], G['exif_shutter'], rtol=1e-15)
  np.testing.assert_allclose(raw_utils.match_images_affine(G['aff_est'], G['aff_gt']), G['aff_out'], rtol=1e-10)
  with pytest.raises(ValueError):
    raw_utils.postprocess_raw(G['pp_raw'][..., :2], G['pp_cam2rgb'])
  # the measured mosaic values survive demosaicking
  d = raw_utils.bilinear_demosaic(G['bayer'])
  assert np.array_equal(d[0::2, 0::2, 0], G['bayer'][0::2, 0::2]) and np.array_equal(d[1::2, 1::2, 2], G['bayer'][1::2, 1::2])


def test_pose_algebra_matches_reference_run():
  from multinerf_b200 import camera_utils as cu
  poses, ring = G['poses'], G['ring']
  rp, rt = cu.recenter_poses(poses)
  np.testing.assert_allclose(rp, G['recenter_poses'], atol=1e-12)
  np.testing.assert_allclose(rt, G['recenter_transform'], atol=1e-12)
  np.testing.assert_allclose(cu.average_pose(poses), G['average_pose'], atol=1e-12)
  np.testing.assert_allclose(cu.pad_poses(poses), G['pad_poses'], atol=0)
  np.testing.assert_allclose(cu.focus_point_fn(ring), G['focus_point'], atol=1e-12)
  pp, pt = cu.transform_poses_pca(ring.copy())
  np.testing.assert_allclose(pp, G['pca_poses'], atol=1e-10)
  np.testing.assert_allclose(pt, G['pca_transform'], atol=1e-10)
  np.testing.assert_allclose(cu.generate_ellipse_path(pp, n_frames=10, z_variation=0.3, z_phase=0.25), G['ellipse'],
                             atol=2e-6)        # the reference resamples theta in float32 (jnp)
  np.testing.assert_allclose(cu.generate_ellipse_path(pp, n_frames=7, const_speed=False), G['ellipse_plain'], atol=1e-12)
  np.testing.assert_allclose(cu.generate_spiral_path(rp, G['bounds'], n_frames=8), G['spiral'], atol=1e-12)
  np.testing.assert_allclose(cu.generate_interpolated_path(ring[:6], n_interp=4), G['interp_path'], atol=1e-9)
  np.testing.assert_allclose(
      cu.interpolate_1d(np.log(np.array([1., 2., 1.5, 3., 2.5, 4., 3.])), 3, 5, 20), G['interp_1d'], atol=1e-12)


# ------------------------------------------------------------------------------------------ datasets
def _write_blender_scene(root, n=3, H=6, W=8, with_normals=False):
  from PIL import Image
  rng = np.random.default_rng(5)
  os.makedirs(os.path.join(root, 'train'), exist_ok=True)
  imgs, frames = [], []
  for i in range(n):
    rgba = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    Image.fromarray(rgba).save(os.path.join(root, 'train', f'r_{i}.png'))
    if with_normals:
      Image.fromarray(rng.integers(0, 256, (H, W, 3), dtype=np.uint8)).save(os.path.join(root, 'train', f'r_{i}_normal.png'))
    m = np.eye(4)
    m[:3, 3] = rng.normal(size=3)
    frames.append({'file_path': f'./train/r_{i}', 'transform_matrix': m.tolist()})
    imgs.append(rgba)
  for split in ('train', 'test'):
    with open(os.path.join(root, f'transforms_{split}.json'), 'w') as f:
      json.dump({'camera_angle_x': 0.7, 'frames': frames}, f)
  return np.stack(imgs), frames


def _no_cast(ds):
  """Host half only (no GPU here): the queue element before `_finish` casts rays on the device."""
  return ds._queue.queue[0]


def test_blender_loader_and_batches(tmp_path):
  from multinerf_b200 import configs, datasets, utils
  imgs, frames = _write_blender_scene(str(tmp_path))
  cfg = configs.Config(dataset_loader='blender', batch_size=16, batching='single_image', near=2., far=6.)
  ds = datasets.load_dataset('train', str(tmp_path), cfg, device='cpu')
  assert ds.size == 3 and (ds.height, ds.width) == (6, 8)
  rgba = imgs.astype(np.float32) / 255.
  want = rgba[..., :3] * rgba[..., 3:] + (1 - rgba[..., 3:])                 # white background
  np.testing.assert_allclose(ds.images, want, atol=1e-6)
  assert ds.focal == pytest.approx(.5 * 8 / np.tan(.35))
  np.testing.assert_allclose(ds.camtoworlds[1], np.array(frames[1]['transform_matrix'], np.float32))
  b = _no_cast(ds)
  px = b.rays
  assert isinstance(px, utils.Pixels) and px.pix_x_int.shape == (16, 1, 1) and b.rgb.shape == (16, 1, 1, 3)
  assert len(np.unique(px.cam_idx)) == 1                                     # single_image batching
  assert px.near.shape == (16, 1, 1, 1) and float(px.near.max()) == 2. and float(px.far.min()) == 6.
  i = 5
  np.testing.assert_array_equal(b.rgb[i, 0, 0], ds.images[px.cam_idx[i, 0, 0, 0], px.pix_y_int[i, 0, 0], px.pix_x_int[i, 0, 0]])
  # test split: one full image per element, in camera order, wrapping around (datasets.py:519-525)
  dt = datasets.load_dataset('test', str(tmp_path), cfg, device='cpu')
  first = _no_cast(dt)
  assert first.rays.pix_x_int.shape == (6, 8) and first.rgb.shape == (6, 8, 3)
  np.testing.assert_array_equal(first.rgb, dt.images[0])
  cfg2 = configs.Config(dataset_loader='blender', batch_size=16, factor=2)
  d2 = datasets.load_dataset('train', str(tmp_path), cfg2, device='cpu')
  assert (d2.height, d2.width) == (3, 4)
  with pytest.raises(ValueError):
    datasets.load_dataset('train', str(tmp_path), configs.Config(dataset_loader='blender', render_path=True), device='cpu')
  with pytest.raises(ValueError):                                             # patch larger than the batch
    datasets.load_dataset('train', str(tmp_path), configs.Config(dataset_loader='blender', batch_size=8, patch_size=4),
                          device='cpu')


def test_patch_and_bayer_batches(tmp_path):
  from multinerf_b200 import configs, datasets, raw_utils
  _write_blender_scene(str(tmp_path), H=12, W=12)
  cfg = configs.Config(dataset_loader='blender', batch_size=32, patch_size=2, apply_bayer_mask=True,
                       num_border_pixels_to_mask=2)
  ds = datasets.load_dataset('train', str(tmp_path), cfg, device='cpu')
  b = _no_cast(ds)
  px = b.rays
  assert px.pix_x_int.shape == (8, 2, 2)                                     # 32 rays = 8 patches of 2 x 2
  assert (px.pix_x_int[:, 0, 1] - px.pix_x_int[:, 0, 0] == 1).all() and (px.pix_y_int[:, 1, 0] - px.pix_y_int[:, 0, 0] == 1).all()
  assert px.pix_x_int.min() >= 2 and px.pix_x_int.max() <= 12 - 2 - 1
  np.testing.assert_array_equal(px.lossmult, raw_utils.pixels_to_bayer_mask(px.pix_x_int, px.pix_y_int))


def _write_colmap(sparse_dir, names, w2c_list, model_id, params, W, H):
  os.makedirs(sparse_dir, exist_ok=True)
  with open(os.path.join(sparse_dir, 'cameras.bin'), 'wb') as f:
    f.write(struct.pack('<Q', 1))
    f.write(struct.pack('<iiQQ', 1, model_id, W, H))
    f.write(struct.pack('<' + 'd' * len(params), *params))
  with open(os.path.join(sparse_dir, 'images.bin'), 'wb') as f:
    f.write(struct.pack('<Q', len(names)))
    for i, (name, (q, t)) in enumerate(zip(names, w2c_list)):
      f.write(struct.pack('<idddddddi', i + 1, *q, *t, 1))
      f.write(name.encode() + b'\x00')
      f.write(struct.pack('<Q', 2))
      f.write(struct.pack('<ddq', 1.0, 2.0, -1) * 2)


def test_llff_loader_with_colmap_model(tmp_path):
  """COLMAP binary model -> poses in the NeRF frame, OPENCV distortion, 360 normalisation, llffhold split."""
  from PIL import Image
  from multinerf_b200 import camera_utils, configs, datasets
  root = str(tmp_path)
  rng = np.random.default_rng(9)
  n, H, W = 9, 8, 12
  names = [f'img_{i:02d}.jpg' for i in range(n)]
  os.makedirs(os.path.join(root, 'images'))
  os.makedirs(os.path.join(root, 'images_2'))
  for nm in names:
    Image.fromarray(rng.integers(0, 256, (H, W, 3), dtype=np.uint8)).save(os.path.join(root, 'images', nm))
    Image.fromarray(rng.integers(0, 256, (H // 2, W // 2, 3), dtype=np.uint8)).save(
        os.path.join(root, 'images_2', nm.replace('.jpg', '.png')))
  w2c = []
  for i in range(n):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w2c.append((q, rng.normal(size=3) * 2))
  _write_colmap(os.path.join(root, 'sparse', '0'), names, w2c, 4, [100., 110., 6., 4., 0.01, -0.02, 0.001, 0.002], W, H)
  cams, images = datasets.read_colmap_model(os.path.join(root, 'sparse', '0'))
  assert cams[1][0] == 'OPENCV' and [im[0] for im in images] == names
  nm, poses, pixtocam, params, camtype = datasets.load_colmap_posedata(os.path.join(root, 'sparse', '0'))
  assert params == {'k1': 0.01, 'k2': -0.02, 'k3': 0., 'p1': 0.001, 'p2': 0.002}
  assert camtype == camera_utils.ProjectionType.PERSPECTIVE
  np.testing.assert_allclose(np.linalg.inv(pixtocam), camera_utils.intrinsic_matrix(100., 110., 6., 4.), atol=1e-9)
  # camera centre = -R^T t; third column = back direction = -(R^T e_z)
  R0 = datasets._qvec_to_rot(w2c[0][0])
  np.testing.assert_allclose(poses[0][:, 3], -R0.T @ w2c[0][1], atol=1e-9)
  np.testing.assert_allclose(poses[0][:, 2], -(R0.T @ np.array([0, 0, 1.])), atol=1e-9)
  assert abs(np.linalg.det(R0) - 1) < 1e-9
  cfg = configs.Config(dataset_loader='llff', factor=2, batch_size=8, near=0.2, far=1e6)
  tr = datasets.load_dataset('train', root, cfg, device='cpu')
  te = datasets.load_dataset('test', root, cfg, device='cpu')
  assert te.size == 2 and tr.size == 7 and (tr.height, tr.width) == (4, 6)          # llffhold = 8: images 0 and 8
  assert np.abs(tr.camtoworlds[:, :3, 3]).max() <= 1.0 + 1e-9                         # PCA-normalised into the unit cube
  np.testing.assert_allclose(tr.pixtocams, (pixtocam @ np.diag([2, 2, 1.])).astype(np.float32))
  assert tr.distortion_params == params and tr.render_poses.shape == (120, 3, 4)
  ff = datasets.load_dataset('train', root, configs.Config(dataset_loader='llff', factor=2, batch_size=8, forward_facing=True),
                             device='cpu')
  assert ff.pixtocam_ndc is not None and ff.cameras[3] is ff.pixtocam_ndc
  rp = datasets.load_dataset('test', root, configs.Config(dataset_loader='llff', factor=2, batch_size=8, render_path=True,
                                                           render_path_frames=5), device='cpu')
  assert rp.size == 5 and _no_cast(rp).rgb is None


def _png(path, arr):
  from PIL import Image
  os.makedirs(os.path.dirname(path), exist_ok=True)
  Image.fromarray(arr).save(path)


def test_tanks_and_temples_nerfpp_loader(tmp_path):
  """NeRF++'s layout (datasets.py:720-764): <split>/{pose,intrinsics,rgb}/*, poses flipped to OpenGL axes,
  focal from the first intrinsics matrix; render_path reads camera_path/ and only the image size from test/rgb."""
  from multinerf_b200 import camera_utils, configs, datasets
  rng = np.random.default_rng(11)
  root = str(tmp_path)
  poses, imgs = [], []
  for split, n in (('train', 3), ('test', 2), ('camera_path', 4)):
    for i in range(n):
      m = np.eye(4)
      m[:3, :3] = np.linalg.qr(rng.normal(size=(3, 3)))[0]
      m[:3, 3] = rng.normal(size=3)
      K = np.eye(4)
      K[0, 0] = K[1, 1] = 37.5
      K[0, 2], K[1, 2] = 5., 3.
      for d, mat in (('pose', m), ('intrinsics', K)):
        os.makedirs(os.path.join(root, split, d), exist_ok=True)
        np.savetxt(os.path.join(root, split, d, f'{i:03d}.txt'), mat.reshape(1, 16))
      if split != 'camera_path':
        img = rng.integers(0, 256, (6, 10, 3), dtype=np.uint8)
        _png(os.path.join(root, split, 'rgb', f'{i:03d}.png'), img)
      if split == 'train':
        poses.append(m)
        imgs.append(img)
  cfg = configs.Config(dataset_loader='tat_nerfpp', batch_size=8)
  ds = datasets.load_dataset('train', root, cfg, device='cpu')
  assert ds.size == 3 and (ds.height, ds.width) == (6, 10) and ds.focal == pytest.approx(37.5)
  np.testing.assert_allclose(ds.images, np.stack(imgs).astype(np.float32) / 255., atol=1e-7)
  np.testing.assert_allclose(ds.camtoworlds, np.stack(poses) @ np.diag([1., -1., -1., 1.]), atol=1e-12)
  np.testing.assert_allclose(ds.pixtocams, camera_utils.get_pixtocam(37.5, 10, 6), atol=1e-12)
  b = _no_cast(ds)
  assert b.rays.pix_x_int.shape == (8, 1, 1) and b.rgb.shape == (8, 1, 1, 3)
  dr = datasets.load_dataset('test', root, configs.Config(dataset_loader='tat_nerfpp', render_path=True), device='cpu')
  assert dr.size == 4 and dr.images is None and (dr.height, dr.width) == (6, 10)


def test_tanks_and_temples_fvs_loader(tmp_path):
  """Free View Synthesis' layout (datasets.py:767-829): dense/ibr3d_pw_<scale>/{im_*.jpg|png, Ks, Rs, ts}.npy; world-to-camera
  [R|t] inverted and flipped, PCA-aligned; every llffhold-th view is the test split; `factor` indexes the size list
  from the largest."""
  from multinerf_b200 import camera_utils, configs, datasets
  rng = np.random.default_rng(12)
  n = 9
  Rs = np.stack([np.linalg.qr(rng.normal(size=(3, 3)))[0] for _ in range(n)])
  ts = rng.normal(size=(n, 3))
  Ks = np.tile(np.array([[50., 0, 4], [0, 50., 3], [0, 0, 1]]), (n, 1, 1))
  imgs = {}
  for scale, (H, W) in (('0.25', (3, 4)), ('0.50', (6, 8))):
    d = os.path.join(str(tmp_path), 'dense', f'ibr3d_pw_{scale}')
    os.makedirs(d)
    imgs[scale] = rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    for i in range(n):
      _png(os.path.join(d, f'im_{i:08d}.png'), imgs[scale][i])
    for name, arr in (('Ks', Ks * (1.0 if scale == '0.50' else 0.5)), ('Rs', Rs), ('ts', ts)):
      np.save(os.path.join(d, f'{name}.npy'), arr)
  cfg = configs.Config(dataset_loader='tat_fvs', batch_size=8, factor=0, llffhold=4)
  ds = datasets.load_dataset('train', str(tmp_path), cfg, device='cpu')
  keep = np.array([i for i in range(n) if i % 4 != 0])
  assert ds.size == len(keep) and (ds.height, ds.width) == (6, 8) and ds.focal == pytest.approx(50.)
  np.testing.assert_allclose(ds.images, imgs['0.50'][keep].astype(np.float32) / 255., atol=1e-7)
  w2c = np.concatenate([Rs, ts[..., None]], -1)
  c2w = np.linalg.inv(camera_utils.pad_poses(w2c))[:, :3, :4] @ np.diag([1., -1., -1., 1.])
  want, _ = camera_utils.transform_poses_pca(c2w)
  np.testing.assert_allclose(ds.camtoworlds, want[keep], atol=1e-10)
  dt = datasets.load_dataset('test', str(tmp_path), cfg, device='cpu')
  assert dt.size == 3
  np.testing.assert_allclose(dt.camtoworlds, want[[0, 4, 8]], atol=1e-10)
  d1 = datasets.load_dataset('train', str(tmp_path), configs.Config(dataset_loader='tat_fvs', factor=1, llffhold=4), device='cpu')
  assert (d1.height, d1.width) == (3, 4) and d1.focal == pytest.approx(25.)
  with pytest.raises(ValueError):
    datasets.load_dataset('train', str(tmp_path), configs.Config(dataset_loader='tat_fvs', factor=2), device='cpu')
  dp = datasets.load_dataset('test', str(tmp_path), configs.Config(dataset_loader='tat_fvs', render_path=True,
                                                                    render_path_frames=6, llffhold=4), device='cpu')
  assert dp.images is None and dp.camtoworlds.shape[0] == 6


def test_dtu_loader(tmp_path):
  """DTU (datasets.py:832-911): rect_XXX_<light>.png under <scan>/, projection matrices under ../../cal18/pos_XXX.txt
  decomposed into intrinsics and pose; poses recentred, scaled into the unit cube, flipped to OpenGL axes."""
  from multinerf_b200 import camera_utils, configs, datasets
  rng = np.random.default_rng(13)
  root = str(tmp_path)
  scan = os.path.join(root, 'Rectified', 'scan1')
  cal = os.path.join(root, 'cal18')
  os.makedirs(scan)
  os.makedirs(cal)
  n, H, W = 9, 6, 8
  K = np.array([[60., 0, 4], [0, 60., 3], [0, 0, 1]])
  imgs, c2ws = [], []
  for i in range(1, n + 1):
    R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    if np.linalg.det(R) < 0:
      R[:, 0] *= -1
    C = rng.normal(size=3) * 2
    P = K @ np.concatenate([R, (-R @ C)[:, None]], -1)          # world -> pixel, camera centre C
    np.savetxt(os.path.join(cal, f'pos_{i:03d}.txt'), P)
    for light in range(7):                                        # 7 light conditions + 'max' = 8 files per view
      img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
      _png(os.path.join(scan, f'rect_{i:03d}_{light}_r5000.png'), img)
      if light == 3:
        imgs.append(img)
    _png(os.path.join(scan, f'rect_{i:03d}_max.png'), rng.integers(0, 256, (H, W, 3), dtype=np.uint8))
    pose = np.eye(4)
    pose[:3, :3] = R.T
    pose[:3, 3] = C
    c2ws.append(pose[:3])
  cfg = configs.Config(dataset_loader='dtu', batch_size=8, factor=0)
  ds = datasets.load_dataset('train', scan, cfg, device='cpu')
  keep = np.array([i for i in range(n) if i % 8 != 0])
  assert ds.size == len(keep) and (ds.height, ds.width) == (H, W)
  np.testing.assert_allclose(ds.images, np.stack(imgs)[keep].astype(np.float32) / 255., atol=1e-7)
  np.testing.assert_allclose(ds.pixtocams[0], np.linalg.inv(K), atol=1e-4)
  want, _ = camera_utils.recenter_poses(np.stack(c2ws).astype(np.float32))
  want = want.copy()
  want[:, :3, -1] /= np.max(np.abs(want[:, :3, -1]))
  want = want @ np.diag([1., -1., -1., 1.]).astype(np.float32)
  np.testing.assert_allclose(ds.camtoworlds, want[keep], atol=2e-4)
  assert np.abs(ds.camtoworlds[:, :3, 3]).max() <= 1.0 + 1e-5
  dt = datasets.load_dataset('test', scan, cfg, device='cpu')
  assert dt.size == 2                                              # views 0 and 8
  with pytest.raises(ValueError):
    datasets.load_dataset('train', scan, configs.Config(dataset_loader='dtu', render_path=True), device='cpu')
