for v in "$@"; do bash profiles/scripts/variant.sh $v; done
