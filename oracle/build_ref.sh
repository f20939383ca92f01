#!/bin/bash
# TEST INFRASTRUCTURE: compiles the UNMODIFIED reference sources for the visibility path where they lie under
# $REF (default /root/reference) into oracle/_ref/libu3dref.so, together with oracle/ref_harness.cpp.
# Nothing is written into the reference tree; no reference source is copied into this repo.
# Recipe = SURVEY.md 8(c).  Canonical float mode: -ffp-contract=off (no FMA contraction), -O3 -msse3.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
SRC="$REF/Source/Urho3D"
TP="$REF/Source/ThirdParty"
JOBS="${JOBS:-$(nproc)}"
if [ ! -d "$SRC" ]; then echo "reference not present at $REF: keeping prebuilt $OUT (if any)"; exit 0; fi
mkdir -p "$OUT/obj" "$OUT/inc" "$OUT/inc2/Urho3D"
# include roots: <Urho3D/...> -> reference sources; <SDL/...>, <PugiXml/...>, <LZ4/...> etc. -> ThirdParty
ln -sfn "$SRC" "$OUT/inc/Urho3D"
ln -sfn "$TP/SDL/include" "$OUT/inc/SDL"
ln -sfn "$TP/PugiXml/src" "$OUT/inc/PugiXml"
CXXFLAGS="-std=c++11 -O3 -DNDEBUG -msse3 -ffp-contract=off -fPIC -w -DGLEW_NO_GLU -DGLEW_STATIC -DURHO3D_STATIC_DEFINE -DURHO3D_THREADING -DURHO3D_LOGGING -DURHO3D_OPENGL -DURHO3D_SSE \
 -I$OUT -I$HERE/ref_stub -I$HERE/ref_stub/Urho3D -I$HERE/ref_stub/sdlcfg -I$OUT/inc -I$SRC -I$REF/Source -I$TP -I$TP/SDL/include -I$TP/rapidjson/include -I$TP/WebP/src"
# The three OctreeQuery subclasses that are file-local in View.cpp (ShadowCasterOctreeQuery, ZoneOccluderOctreeQuery,
# OccludedFrustumOctreeQuery, View.cpp:58-157) are compiled from the reference's own text: the lines between the first class comment and
# CheckVisibilityWork are extracted into a generated include (git-ignored, under oracle/_ref/) that ref_harness.cpp pulls into namespace
# Urho3D.  View.cpp itself needs the GL back-end and cannot be compiled here.
VIEWCPP="$SRC/Graphics/View.cpp"
awk '/^\/\/\/ %Frustum octree query for shadowcasters\./{on=1} /^void CheckVisibilityWork\(/{on=0} on' "$VIEWCPP" > "$OUT/view_queries.inc"
for c in "class ShadowCasterOctreeQuery" "class ZoneOccluderOctreeQuery" "class OccludedFrustumOctreeQuery"; do
  grep -q "$c" "$OUT/view_queries.inc" || { echo "could not extract '$c' from $VIEWCPP"; exit 1; }
done
# Two member functions of View that depend only on classes compiled here are also taken from the reference's own text and compiled as
# members of a stand-in class (ref_harness.cpp: RefViewShim, `#define View RefViewShim` around the include): View::UpdateOccluders
# (View.cpp:2171-2229; touches renderer_->GetOccluderSizeThreshold() and frame_) and View::IsShadowCasterVisible (View.cpp:2455-2489;
# touches frame_), View::DrawOccluders (View.cpp:2231-2275; maxOccluderTriangles_, activeOccluders_), View::ProcessShadowCasters (View.cpp:2379-2453; its Light and LightQueryResult are stand-ins with the members it reads), and
# the free function CheckVisibilityWork (View.cpp:160-227; reads occlusionBuffer_, cullCamera_, cameraZoneOverride_,
# sceneResults_, frame_ and calls FindZone through its View pointer).
awk '/^void View::UpdateOccluders\(/{on=1} /^void View::DrawOccluders\(/{on=0} on' "$VIEWCPP" > "$OUT/view_update_occluders.inc"
awk '/^bool View::IsShadowCasterVisible\(/{on=1} /^IntRect View::GetShadowMapViewport\(/{on=0} on' "$VIEWCPP" > "$OUT/view_shadow_caster_visible.inc"
awk '/^void View::DrawOccluders\(/{on=1} /^void View::ProcessLight\(/{on=0} on' "$VIEWCPP" > "$OUT/view_draw_occluders.inc"
grep -q "bool success = occluder->DrawOcclusion(buffer)" "$OUT/view_draw_occluders.inc" || { echo "could not extract View::DrawOccluders from $VIEWCPP"; exit 1; }
awk '/^void View::ProcessShadowCasters\(/{on=1} /^bool View::IsShadowCasterVisible\(/{on=0} on' "$VIEWCPP" > "$OUT/view_process_shadow_casters.inc"
grep -q "query.shadowCasters_.Push(drawable)" "$OUT/view_process_shadow_casters.inc" || { echo "could not extract View::ProcessShadowCasters from $VIEWCPP"; exit 1; }
awk '/^void CheckVisibilityWork\(/{on=1} /^void ProcessLightWork\(/{on=0} on' "$VIEWCPP" > "$OUT/view_check_visibility_work.inc"
grep -q "drawable->MarkInView(view->frame_)" "$OUT/view_check_visibility_work.inc" || { echo "could not extract CheckVisibilityWork from $VIEWCPP"; exit 1; }
grep -q "Sort(occluders.Begin(), occluders.End(), CompareDrawables)" "$OUT/view_update_occluders.inc" || { echo "could not extract View::UpdateOccluders from $VIEWCPP"; exit 1; }
grep -q "lightViewFrustum.IsInsideFast(lightViewBox)" "$OUT/view_shadow_caster_visible.inc" || { echo "could not extract View::IsShadowCasterVisible from $VIEWCPP"; exit 1; }
FILES=""
for d in Container Math Core IO Resource Scene; do FILES="$FILES $(ls $SRC/$d/*.cpp)"; done
for f in OcclusionBuffer Octree OctreeQuery Camera Drawable GraphicsDefs; do FILES="$FILES $SRC/Graphics/$f.cpp"; done
FILES="$FILES $TP/PugiXml/src/pugixml.cpp $HERE/ref_harness.cpp $HERE/ref_stubs.cpp"
MK="$OUT/obj/build.mk"
: > "$MK"
OBJS=""
for f in $FILES; do
  o="$OUT/obj/$(echo "$f" | sed -e "s#$REF/##" -e "s#$HERE/##" -e 's#[/.]#_#g').o"
  OBJS="$OBJS $o"
  printf '%s: %s\n\tg++ %s -c %s -o %s\n' "$o" "$f" "$CXXFLAGS" "$f" "$o" >> "$MK"
done
for f in lz4 lz4hc; do
  o="$OUT/obj/tp_$f.o"; OBJS="$OBJS $o"
  printf '%s: %s\n\tgcc -O2 -fPIC -w -c %s -o %s\n' "$o" "$TP/LZ4/$f.c" "$TP/LZ4/$f.c" "$o" >> "$MK"
done
printf 'all:%s\n' "$OBJS" >> "$MK"
make -s -j"$JOBS" -f "$MK" all
g++ -shared -o "$OUT/libu3dref.so" $OBJS -lpthread -ldl -Wl,-z,defs 2> "$OUT/link.log" || { sort -u "$OUT/link.log" | grep -o "undefined reference to .*" | sort -u | head -80; exit 1; }
echo "built $OUT/libu3dref.so"
