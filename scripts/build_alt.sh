# Builds an alternative copy of the library with extra nvcc flags into adaprompt_b200/_alt/<name>.so (git-ignored; used for
# A/B measurements on the GPU box).  Usage: build_alt.sh <name> <flags...>
name=$1; shift
mkdir -p adaprompt_b200/_alt /tmp/alt_$name
objs=""
for f in adaprompt_b200/csrc/*.cu; do
  o=/tmp/alt_$name/$(basename $f .cu).o
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -I include "$@" -c $f -o $o &
  objs="$objs $o"
done
wait
nvcc -shared -o adaprompt_b200/_alt/$name.so $objs -cudart static -gencode arch=compute_100a,code=sm_100a && ls -la adaprompt_b200/_alt/$name.so
