"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

CPU restatement of the mask -> polygon step that follows the segmentation path on the OMERO route
(SURVEY.md 8(f) N2):

  * ``get_indices_pandas``  /root/reference/src/utils/hull_polygon.py:8-42   pixel positions per mask id
  * ``cv2_countour``        /root/reference/src/utils/hull_polygon.py:45-89  per id: crop with a 1-px frame,
                            ``cv2.findContours(RETR_TREE, CHAIN_APPROX_NONE)``, (x,y) -> (y,x), shift back
  * caller                  /root/reference/src/inference/infer.py:273-289   "x,y " strings per polygon

PARITY UNPINNED: OpenCV (cv2 4.5.5, requirements.yml) is neither vendored nor installable here and the reference has
no fixtures for this step.  ``find_outer_contour`` restates OpenCV's border following (modules/imgproc/src/contours.cpp,
``icvFetchContour`` with CV_CHAIN_APPROX_NONE on an outer border: Suzuki & Abe 1985, 8-connectivity): start at the first
raster pixel of the component, find the last border pixel i1 by a clockwise search that starts at the west neighbour,
then follow the border with counter-clockwise searches, emitting every visited pixel, until the step i1 -> i0 comes
up again.  Scope: every mask id is one 8-connected component (always true for watershed labels below 65536); for a
component with holes the reference ends up returning the outer contour (hull_polygon.py:61-75), which is what this
returns.
"""
import numpy as np

# OpenCV's direction codes (x, y): 0 E, 1 NE, 2 N, 3 NW, 4 W, 5 SW, 6 S, 7 SE  (y grows downwards)
DX = (1, 1, 0, -1, -1, -1, 0, 1)
DY = (0, -1, -1, -1, 0, 1, 1, 1)


def find_outer_contour(img):
    """``img``: 2-D array whose non-zero pixels are one component, with an all-zero 1-px frame.
    Returns the (N, 2) int array of (x, y) points, as cv2.findContours(..., CHAIN_APPROX_NONE)[0][0].squeeze()."""
    fg = np.asarray(img) != 0
    ys, xs = np.nonzero(fg)
    if len(ys) == 0:
        return np.zeros((0, 2), np.int64)
    k = np.lexsort((xs, ys))[0]              # first pixel in raster order
    x0, y0 = int(xs[k]), int(ys[k])
    s_end = s = 4
    while True:
        s = (s - 1) & 7
        x1, y1 = x0 + DX[s], y0 + DY[s]
        if fg[y1, x1] or s == s_end:
            break
    if s == s_end and not fg[y0 + DY[s], x0 + DX[s]]:
        return np.array([[x0, y0]], np.int64)            # single pixel
    pts = []
    x3, y3 = x0, y0
    while True:
        s_end = s
        while True:
            s += 1
            x4, y4 = x3 + DX[s & 7], y3 + DY[s & 7]
            if fg[y4, x4]:
                break
        s &= 7
        pts.append((x3, y3))
        if (x4, y4) == (x0, y0) and (x3, y3) == (x1, y1):
            break
        x3, y3 = x4, y4
        s = (s + 4) & 7
    return np.array(pts, np.int64)


def cv2_contour_of_label(mask, label):
    """hull_polygon.py:45-89 for one mask id (single-contour branch :83-89): [(2, N) array, rows = (y, x)]."""
    ys, xs = np.nonzero(np.asarray(mask) == label)
    mn_y, mn_x = ys.min(), xs.min()
    img = np.zeros((ys.max() - mn_y + 3, xs.max() - mn_x + 3), np.uint8)
    img[ys - mn_y + 1, xs - mn_x + 1] = 255
    pts = find_outer_contour(img)[:, ::-1]                   # (x, y) -> (y, x)
    pts = pts + np.array([[mn_y, mn_x]]) - 1
    return [pts.T]


def mask_to_polygons(mask):
    """{mask id: [(2, N) (y; x) arrays]} for every non-zero id, ids ascending (get_indices_pandas groups sorted)."""
    mask = np.asarray(mask)
    return {int(i): cv2_contour_of_label(mask, i) for i in np.unique(mask[mask != 0])}


def points_string(polygon_points):
    """infer.py:281-284: "x,y " for every contour point"""
    return "".join("{},{} ".format(polygon_points[1, c], polygon_points[0, c]) for c in range(polygon_points.shape[1]))
