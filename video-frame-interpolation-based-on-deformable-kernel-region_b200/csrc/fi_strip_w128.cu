// fi_strip_w128.cu -- the "_ori" strip kernel of fi_strip.cu once more with 128-column tiles (16 compute warps, four filter
// stages): the instantiation for widths where 144-column strips would leave a mostly empty last strip.
#define VFIDKR_ORI_TW 128
#define VFIDKR_ORI_ENTRY fi_strip_forward_ori_w128
#define VFIDKR_STRIP_NS strip_w128
#include "fi_strip.cu"
