"""``Slice3DDataset`` -- drop-in for the reference dataset of the same name (reference: reg_slices/src/datasets.py:14-177;
SURVEY.md section 8 row f-4), reading the same on-disk layout::

    <dir_data>/<name_dataset>/03_splits/<category>/<split>.lst      shape ids
                              00_img_input/<shape>/<view:03d>.png   rendered input views (RGBA)
                              00_img_input/<shape>/meta.pkl         [_, az[], el[], distance[], _, scale, offset]
                              01_img_slices/<shape>/<view>/X_1.png  the 12 slice images (RGBA; 04_img_slices_gen /
                                                                    05_img_slices_rec: RGB, already at img_size)
                              02_sdfs/<shape>.npy                   (n, 4) xyz + sdf at the 0.003 level set

Same constructor (``split``, ``args``), ``__len__`` and ``__getitem__ -> feed_dict`` (keys, shapes, dtypes and VALUES:
the image half uses the integer resample tables of ``slice3d_b200.inputs`` and equals PIL + torchvision bit for bit; the
test split's view and query subset are the reference's -- view 4, numpy legacy seed 1234).  So it works under a
``DataLoader`` exactly like the reference's.

What it adds is ``batch(indices, device)``: the samples' PNGs are only DECODED on the host (inflate is byte-serial work);
compositing, resize, to-tensor and normalise of all 13 x B images run as ONE ``inputs.preprocess_rgba`` call on the GPU
(two kernels), and the batched feed_dict is assembled on the device -- the training loop needs no worker processes for
the per-pixel work.  A CPU ``device`` uses the host mirror of the same arithmetic.
"""
import os
import pickle
import random

import numpy as np
import torch
from torch.utils.data import Dataset

from . import inputs


class Slice3DDataset(Dataset):
    def __init__(self, split, args):
        # datasets.py:15-53
        self.split = split
        self.n_qry = args.n_qry
        self.dir_dataset = os.path.join(args.dir_data, args.name_dataset)
        self.name_dataset = args.name_dataset
        self.img_size = args.img_size
        self.files = []
        if self.name_dataset == "shapenet":
            categories = (args.categories_train if split in {"train", "val"} else args.categories_test).split(",")[:-1]
        else:
            categories = [""]
        for category in categories:
            with open(f"{self.dir_dataset}/03_splits/{category}/{split}.lst") as f:
                self.files += [(category, shape_id) for shape_id in f.read().split()]
        self.dir_sfd = f"{self.dir_dataset}/02_sdfs/"
        self.from_which_slices = args.from_which_slices
        self.dir_img_slice = {"gen": f"{self.dir_dataset}/04_img_slices_gen", "gt": f"{self.dir_dataset}/01_img_slices",
                              "gt_rec": f"{self.dir_dataset}/05_img_slices_rec"}[self.from_which_slices]
        self.dir_img_ipt = f"{self.dir_dataset}/00_img_input"
        self.camera_metainfo = f"{self.dir_dataset}/00_img_input"
        self.use_white_bg = bool(args.use_white_bg)
        self.n_views = args.n_views

    def __len__(self):
        return len(self.files)

    # ------------------------------------------------------------------ host half: file reads and a dozen numbers
    def load_raw(self, index, rng=None):
        """Everything of one sample that is NOT per-pixel work: the decoded images (uint8 arrays, untouched), the camera
        numbers of the chosen view, and the query subset (datasets.py:89-167)."""
        from PIL import Image
        _, shape_id = self.files[index]
        view = random.randint(0, self.n_views - 1) if self.split == "train" else 4  # datasets.py:92-95
        cmr = "%03d" % view
        img_ipt = np.array(Image.open(f"{self.dir_img_ipt}/{shape_id}/{cmr}.png"))
        slices = [np.array(Image.open(f"{self.dir_img_slice}/{shape_id}/{cmr}/{stem}.png")) for stem in inputs.SLICE_ORDER]
        with open(f"{self.camera_metainfo}/{shape_id}/meta.pkl", "rb") as f:
            meta = pickle.load(f)
        az, el, distance = -meta[1][view], meta[2][view], meta[3][view]
        sdf_npy = np.load(f"{self.dir_sfd}/{shape_id}.npy")
        qry, occ, sdf = inputs.prepare_queries(sdf_npy, meta[5], meta[6], self.n_qry, self.split, rng)
        return {"img_input": img_ipt, "slices": slices, "az": az, "el": el, "distance": distance,
                "qry": qry, "occ": occ, "sdf": sdf}

    def _slices_ready(self):
        # 'gen' / 'gt_rec' slices are network outputs stored at img_size: ToTensor + Normalize only (datasets.py:42,47,111)
        return self.from_which_slices in ("gen", "gt_rec")

    @staticmethod
    def _to_tensor_normalise(u8):
        """T.ToTensor + T.Normalize(0.5, 0.5) of (N,H,W,C) uint8 -> (N,C,H,W) float32."""
        return u8.permute(0, 3, 1, 2).to(torch.float32).div(255).sub(0.5).div(0.5)

    def _images(self, raws, device):
        """(B,3,S,S) input views and (B,36,S,S) slices of the samples ``raws`` on ``device``."""
        device = torch.device(device)
        S, B = self.img_size, len(raws)
        groups = [[r["img_input"] for r in raws]]
        if not self._slices_ready():
            groups.append([s for r in raws for s in r["slices"]])
        outs = []
        for imgs in groups:
            # one preprocess call per image SHAPE (a dataset's views and slices each come in one size)
            res = [None] * len(imgs)
            shapes = {}
            for i, a in enumerate(imgs):
                if a.ndim != 3 or a.shape[2] != 4:
                    raise ValueError(f"expected an RGBA image, got an array of shape {a.shape} (the reference indexes "
                                     "channel 3 as alpha, datasets.py:75-88)")
                shapes.setdefault(a.shape, []).append(i)
            for idxs in shapes.values():
                stack = torch.from_numpy(np.stack([imgs[i] for i in idxs]))
                if device.type == "cuda":
                    t = inputs.preprocess_rgba(stack.to(device, non_blocking=True), S, self.use_white_bg)
                else:
                    t = inputs.preprocess_rgba_host(stack.numpy(), S, self.use_white_bg)
                for j, i in enumerate(idxs):
                    res[i] = t[j]
            outs.append(torch.stack(res))
        img_input = outs[0]
        if self._slices_ready():
            u8 = torch.from_numpy(np.stack([s for r in raws for s in r["slices"]])).to(device)
            sl = self._to_tensor_normalise(u8)
        else:
            sl = outs[1]
        return img_input, sl.reshape(B, 12 * sl.shape[1], sl.shape[2], sl.shape[3])

    # ------------------------------------------------------------------ reference API
    def __getitem__(self, index):
        """datasets.py:89-177: the feed_dict of one sample (host tensors)."""
        raw = self.load_raw(index)
        img_input, img_slices = self._images([raw], "cpu")
        rot, T = inputs.camera_matrices(raw["az"], raw["el"], raw["distance"])
        return {"img_input": img_input[0], "qry_norot": raw["qry"], "obj_rot_mat": rot, "trans_mat_wo_rot_tp": T,
                "occ": raw["occ"], "sdf": raw["sdf"], "img_slices": img_slices[0]}

    # ------------------------------------------------------------------ batched, per-pixel work on the device
    def batch(self, indices, device, rng=None):
        """The collated feed_dict of the samples ``indices`` on ``device`` (what ``DataLoader`` + ``.cuda()`` deliver in
        train.py:41-42), with the image half of all samples done in one library call on the GPU."""
        raws = [self.load_raw(i, rng) for i in indices]
        img_input, img_slices = self._images(raws, device)
        cams = [inputs.camera_matrices(r["az"], r["el"], r["distance"]) for r in raws]
        dev = torch.device(device)
        put = lambda ts: torch.stack(ts).to(dev, non_blocking=True)  # noqa: E731
        return {"img_input": img_input, "qry_norot": put([r["qry"] for r in raws]), "obj_rot_mat": put([c[0] for c in cams]),
                "trans_mat_wo_rot_tp": put([c[1] for c in cams]), "occ": put([r["occ"] for r in raws]),
                "sdf": put([r["sdf"] for r in raws]), "img_slices": img_slices}
