"""Reference config dict -> flat oracle config.  TEST INFRASTRUCTURE.

Restates the constructor rules of losses/DenseContrastiveLossV2.py:12-31 and
losses/DenseContrastiveLossV2_ms.py:13-31 (defaults, the cross-scale temperature quirk).
``num_all_classes`` comes from the caller (4 integers per dataset).
"""


def oracle_cfg(loss_cfg, num_all_classes):
    c = loss_cfg
    out = dict(
        num_all_classes=num_all_classes,
        temperature=c.get("temperature", 0.5),                       # V2.py:19
        min_views=c.get("min_views_per_class", 5),                    # V2.py:21
        max_views=c.get("max_views_per_class", 2500),                 # V2.py:27
        max_total=c.get("max_features_total", 10000),                 # V2.py:28
        cross_scale=c.get("cross_scale_contrast", False),             # _ms.py:27
        detach_deepest=c.get("detach_deepest", False),                # _ms.py:29
        w_high_low=c.get("w_high_low", 1.0), w_high_mid=c.get("w_high_mid", 1.0),
    )
    scales = c.get("scales", 2)
    out["scales"] = scales
    out["weights"] = c.get("weights", [1.0] * scales)
    # _ms.py:28: temperature unless the key cross_scale_temperature exists, then hard-coded 0.1
    out["cs_temperature"] = c["temperature"] if "cross_scale_temperature" not in c else 0.1
    return out
