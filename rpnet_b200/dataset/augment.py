"""Host-side augmentations of the reference's `mode='train'` readers (dataset/few_shot_reader.py:27-60, 201-228;
dataset/brain_reader.py:208-293), restated.  They run on the CPU inside Dataset.__getitem__ like the reference's and draw from the
same generators in the same order (module `random`, `np.random`, torch's default generator through torchvision), so that a seeded
run reproduces the reference's item bit for bit — except the elastic deformation, which the reference seeds from OS entropy
(`np.random.RandomState(None)`, brain_reader.py:256-257): `elastic_transform` takes the generator explicitly and is pinned that way.

The reference passes `fillcolor=None` to `transforms.RandomAffine` (few_shot_reader.py:29,52), a keyword torchvision removed; its
meaning (constant fill 0 outside the image, nearest interpolation) is the current default, which is what is requested here."""
import numpy as np
import torch


def _random_affine(t, degrees, translate, scale, shear):
    import torchvision.transforms as transforms
    return transforms.RandomAffine(degrees, translate=translate, scale=scale, shear=shear, fill=0)(t)


def random_transform(images, labels):
    """few_shot_reader.py:27-47 — one random affine (±5°, ±20 % shift, zoom 0.7 … 1.5) applied to the query slice and its label
    together; the image is mapped to [0, 1] first and the zeros the warp brings in are replaced by the image minimum."""
    images = (images + 1) / 2
    floor = images.min()
    both = _random_affine(torch.cat([images, labels[None, ...]], dim=1), 5, (0.2, 0.2), (0.7, 1.5), 0)
    images, labels = both[:, [0], ...], both[:, 1, ...]
    images[images == 0] = floor                      # torchvision has no custom padding value for tensors (:44)
    return images * 2 - 1, labels


def random_label_transform(labels):
    """few_shot_reader.py:50-60 — a small random affine of a label map (±5°, ±2 % shift, zoom 0.5 … 1.5, shear ±5°)."""
    return _random_affine(labels[None, None, ...], 5, (0.02, 0.02), (0.5, 1.5), 5)[:, 0, ...]


def _gamma(img, gamma_range):
    g = np.random.rand() * (gamma_range[1] - gamma_range[0]) + gamma_range[0]
    lo = img.min()
    span = (img.max() - lo + 1e-5)
    img = img - lo + 1e-5
    return span * np.power(img * 1.0 / span, g) + lo


def gamma_tansform(img, gamma_range):
    """few_shot_reader.py:201-211 (the reference's spelling) — random gamma on the [0, 1]-mapped slice, back to [-1, 1]."""
    return _gamma((img + 1) / 2., gamma_range) * 2 - 1


def gamma_tansform_with_label(img, label, gamma_range):
    """few_shot_reader.py:214-228 — the same, applied inside the label region only."""
    return img * (1 - label) + gamma_tansform(img, gamma_range) * label


def elastic_transform(image, mask, alpha=1000, sigma=30, alpha_affine=0.04, padding_value=-1., random_state=None):
    """brain_reader.py:248-293 — one in-plane affine jitter (three control points moved by ±alpha_affine) followed by a smooth random
    displacement field (uniform noise, Gaussian sigma, amplitude alpha), the same for every slice; image [1, D, H, W] bilinear with
    `padding_value` outside, mask [C, D, H, W] nearest."""
    import cv2
    from scipy.ndimage import gaussian_filter, map_coordinates
    rs = np.random.RandomState(None) if random_state is None else random_state
    plane = image.shape[2:]
    n_cls, depth, ny, nx = mask.shape
    centre = np.float32(plane) // 2
    half = min(plane) // 3
    src = np.float32([centre + half, [centre[0] + half, centre[1] - half], centre - half])
    dst = src + rs.uniform(-alpha_affine, alpha_affine, size=src.shape).astype(np.float32)
    M = cv2.getAffineTransform(src, dst)
    dx = gaussian_filter((rs.rand(*plane) * 2 - 1), sigma) * alpha
    dy = gaussian_filter((rs.rand(*plane) * 2 - 1), sigma) * alpha
    gx, gy = np.meshgrid(np.arange(nx), np.arange(ny))
    where = np.reshape(gy + dy, (-1, 1)), np.reshape(gx + dx, (-1, 1))
    out_img, out_mask = np.zeros_like(image), np.zeros_like(mask)
    for z in range(depth):
        warped = cv2.warpAffine(image[0, z], M, plane[::-1], borderMode=cv2.BORDER_CONSTANT, borderValue=padding_value)
        out_img[0, z] = map_coordinates(warped, where, order=1, mode='constant', cval=padding_value).reshape(plane)
        for c in range(n_cls):
            if np.any(mask[c, z]):
                # BORDER_TRANSPARENT leaves destination pixels without a source untouched; the reference passes no destination, so
                # OpenCV hands it uninitialised memory there (its label maps can carry garbage in the strip the warp uncovers).  A
                # zeroed destination is what `borderValue=0` asks for and what the reference gets whenever that memory is clean.
                m = cv2.warpAffine(mask[c, z], M, plane[::-1], dst=np.zeros(plane, mask.dtype), flags=cv2.INTER_NEAREST,
                                   borderMode=cv2.BORDER_TRANSPARENT, borderValue=0)
                out_mask[c, z] = map_coordinates(m, where, order=0, mode='constant').reshape(plane)
    return out_img, out_mask


def elastic_transform_all(image, mask, alpha=1000, sigma=30, alpha_affine=0.04, padding_value=-1., random_state=None):
    """brain_reader.py:208-245 — only the in-plane (xy) transform is active in the reference; it seeds it from OS entropy."""
    return elastic_transform(image, mask, alpha=alpha, sigma=sigma, alpha_affine=alpha_affine, padding_value=padding_value,
                             random_state=random_state)
