# usage: bash tools/ab.sh <tool.py> <variant> [variant...]   ("base" = the in-tree libkzd.so)
tool=$1; shift
for v in "$@"; do
  if [ "$v" = base ]; then unset KZD_LIB_PATH; else export KZD_LIB_PATH=$PWD/kuzushiji_vision_b200/variants/libkzd_$v.so; fi
  python $tool "variant=$v" 2>&1 | grep variant
done
