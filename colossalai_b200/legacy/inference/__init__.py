"""Legacy inference tree (reference `colossalai/legacy/inference`): the quantisation back ends live on as
`colossalai_b200.quantization.{gptq,smoothquant}`; serving itself is `colossalai_b200.inference`."""
