function SparseDynamicG1!(T::Vector{<: Real}, g1_v::Vector{<: Real}, y::Vector{<: Real}, x::Vector{<: Real}, params::Vector{<: Real}, steady_state::Vector{<: Real})
    @assert length(T) >= 11
    @assert length(g1_v) == 31
    @assert length(y) == 24
    @assert length(x) == 3
@inbounds begin
g1_v[1]=100;
g1_v[2]=-params[2];
g1_v[3]=-params[6];
g1_v[4]=-params[5];
g1_v[5]=1;
g1_v[6]=-params[7];
g1_v[7]=-T[10];
g1_v[8]=-100;
g1_v[9]=T[2];
g1_v[10]=1;
g1_v[11]=-400;
g1_v[12]=1;
g1_v[13]=-T[6]*params[3];
g1_v[14]=-400;
g1_v[15]=1;
g1_v[16]=-100;
g1_v[17]=-1;
g1_v[18]=params[7];
g1_v[19]=T[10];
g1_v[20]=1;
g1_v[21]=1;
g1_v[22]=1;
g1_v[23]=1;
g1_v[24]=-1;
g1_v[25]=T[11];
g1_v[26]=-T[5];
g1_v[27]=T[11];
g1_v[28]=1;
g1_v[29]=-T[8];
g1_v[30]=-T[7];
g1_v[31]=-T[9];
end
    return nothing
end
