// cvshim: src/pf2DRao.h includes <ros/ros.h> but uses nothing from it
