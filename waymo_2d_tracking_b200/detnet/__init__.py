"""Drop-in mirror of the reference's ``detnet`` package for the post-detection path only:
``detnet.ensemble`` (CLI + functions), ``detnet.nn.tta.nms_detections`` and
``detnet.utils.box_utils.{point_form, center_size, nms}``.  The detector zoo, trainer and
dataset code of the reference's ``detnet`` are out of scope."""
